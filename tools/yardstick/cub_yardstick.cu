// cub_yardstick.cu -- YARDSTICK ONLY (tools/microbench.py --what cub): cub::DeviceRadixSort on the same keys as
// ukm_sort_u64, to say how far the hand-written onesweep (unikmer_b200/csrc/radix_sort.cu) is from NVIDIA's tuned library
// sort on this GPU.  Never linked into libukm.so, never on the product path.
#include <cub/cub.cuh>
#include <stdint.h>

extern "C" int cub_sort_u64(const uint64_t* d_in, uint64_t* d_out, size_t n, int begin_bit, int end_bit, void* stream,
                            void* d_temp, size_t temp_bytes, size_t* temp_needed) {
    size_t need = 0;
    cudaError_t e = cub::DeviceRadixSort::SortKeys(nullptr, need, d_in, d_out, n, begin_bit, end_bit, (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    if (temp_needed) *temp_needed = need;
    if (!d_temp) return 0;
    if (temp_bytes < need) return -1;
    e = cub::DeviceRadixSort::SortKeys(d_temp, need, d_in, d_out, n, begin_bit, end_bit, (cudaStream_t)stream);
    return (int)e;
}
