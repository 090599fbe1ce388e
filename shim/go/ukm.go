// Package ukm is the cgo shim a unikmer maintainer adds to call libukm.so (include/ukm.h)
// from unikmer/cmd/*.go in place of the per-k-mer inner loops.
//
// SOURCE ONLY: this image has no Go toolchain (SURVEY.md F3), so this file is not compiled or
// tested here; everything it binds is exercised through the same C ABI from Python (ctypes) in
// tests/.  Build the reference with CGO_ENABLED=1 and
//   CGO_CFLAGS=-I<repo>/include  CGO_LDFLAGS="-L<repo>/unikmer_b200 -lukm"
// (unikmer/packaging.sh:3 sets CGO_ENABLED=0 today).
//
// cgo rules observed: Go memory is passed for the duration of a call only, the slices contain no
// Go pointers ([]uint64, []uint32, []CodeTaxid), and libukm never retains a host pointer.
package ukm

/*
#cgo LDFLAGS: -lukm
#include <stdlib.h>
#include "ukm.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"unsafe"
)

// Flags mirror UKM_F_* in ukm.h.
const (
	FTaxid        = uint(C.UKM_F_TAXID)
	FMixTaxid     = uint(C.UKM_F_MIX_TAXID)
	FCompareTaxid = uint(C.UKM_F_COMPARE_TAXID)
	FCanonical    = uint(C.UKM_F_CANONICAL)
	FHashed       = uint(C.UKM_F_HASHED)
	FCircular     = uint(C.UKM_F_CIRCULAR)
	FScaled       = uint(C.UKM_F_SCALED)
	FValidate     = uint(C.UKM_F_VALIDATE)
	FShard        = uint(C.UKM_F_SHARD)
)

// Operations of SetopsStream mirror ukm_setop.
const (
	OpInter = int(C.UKM_OP_INTER)
	OpDiff  = int(C.UKM_OP_DIFF)
	OpUnion = int(C.UKM_OP_UNION)
)

// Fold modes mirror ukm_fold_mode (sort.go:482-573, util-sort.go:35-190).
const (
	FoldPlain         = int(C.UKM_FOLD_PLAIN)
	FoldUnique        = int(C.UKM_FOLD_UNIQUE)
	FoldRepeatedFinal = int(C.UKM_FOLD_REPEATED_FINAL)
	FoldRepeatedChunk = int(C.UKM_FOLD_REPEATED_CHUNK)
)

// Ctx owns one GPU (one per process, like one unikmer command).
type Ctx struct{ c *C.ukm_ctx }

// New creates a context on `device`.
func New(device int) (*Ctx, error) {
	c := C.ukm_create(C.int(device))
	if c == nil {
		return nil, fmt.Errorf("ukm: %s", C.GoString(C.ukm_last_error(nil)))
	}
	x := &Ctx{c}
	runtime.SetFinalizer(x, func(x *Ctx) { x.Close() })
	return x, nil
}

// Close releases the context.
func (x *Ctx) Close() {
	if x.c != nil {
		C.ukm_destroy(x.c)
		x.c = nil
	}
}

func (x *Ctx) err(status C.int) error {
	if status == C.UKM_OK {
		return nil
	}
	// unikmer's convention is checkError(err) -> log + os.Exit(-1) (util-cli.go:39-44)
	return fmt.Errorf("ukm (%d): %s", int(status), C.GoString(C.ukm_last_error(x.c)))
}

// Set is one k-mer stream: what unik.Reader.ReadCodeWithTaxid yields for a file.
type Set struct {
	Codes       []uint64
	Taxids      []uint32 // nil: use GlobalTaxid
	GlobalTaxid uint32
	Sorted      bool
}

func span(s *Set) C.ukm_span {
	var sp C.ukm_span
	if len(s.Codes) > 0 {
		sp.keys = (*C.uint64_t)(unsafe.Pointer(&s.Codes[0]))
	}
	if len(s.Taxids) > 0 {
		sp.taxids = (*C.uint32_t)(unsafe.Pointer(&s.Taxids[0]))
	}
	sp.global_taxid = C.uint32_t(s.GlobalTaxid)
	sp.n = C.size_t(len(s.Codes))
	sp.cap = C.size_t(cap(s.Codes))
	sp.where = C.UKM_HOST
	if s.Sorted {
		sp.sorted = 1
	}
	return sp
}

// cgo forbids passing Go memory that itself holds Go pointers, so the span array lives in C memory.
func spans(sets []Set) (*C.ukm_span, func()) {
	n := len(sets)
	p := (*C.ukm_span)(C.malloc(C.size_t(n) * C.size_t(unsafe.Sizeof(C.ukm_span{}))))
	arr := unsafe.Slice(p, n)
	for i := range sets {
		arr[i] = span(&sets[i])
	}
	return p, func() { C.free(unsafe.Pointer(p)) }
}

func outSpan(codes []uint64, taxids []uint32) C.ukm_span {
	var sp C.ukm_span
	if cap(codes) > 0 {
		sp.keys = (*C.uint64_t)(unsafe.Pointer(&codes[:1][0]))
	}
	if cap(taxids) > 0 {
		sp.taxids = (*C.uint32_t)(unsafe.Pointer(&taxids[:1][0]))
	}
	sp.cap = C.size_t(cap(codes))
	sp.where = C.UKM_HOST
	return sp
}

// SetTaxonomy replaces loadTaxonomy's in-memory tree (util.go:119-171): parent[t] (0 = unknown,
// root = itself) from nodes.dmp and old->new pairs from merged.dmp.
func (x *Ctx) SetTaxonomy(parent, mergedFrom, mergedTo []uint32) error {
	var mf, mt *C.uint32_t
	if len(mergedFrom) > 0 {
		mf = (*C.uint32_t)(unsafe.Pointer(&mergedFrom[0]))
		mt = (*C.uint32_t)(unsafe.Pointer(&mergedTo[0]))
	}
	return x.err(C.ukm_set_taxonomy(x.c, (*C.uint32_t)(unsafe.Pointer(&parent[0])), C.size_t(len(parent)), mf, mt, C.size_t(len(mergedFrom))))
}

// SortUint64s replaces sortutil.Uint64s(m): sort.go:274,337,463; union.go:274,295; diff.go:587;
// common.go:344; count.go:581; split.go:311,383.
func (x *Ctx) SortUint64s(m []uint64, keyBits int) error {
	if len(m) < 2 {
		return nil
	}
	return x.err(C.ukm_sort_u64(x.c, (*C.uint64_t)(unsafe.Pointer(&m[0])), C.size_t(len(m)), C.int(keyBits), C.UKM_HOST))
}

// CodeTaxid has the layout of cmd.CodeTaxid (kmers.go:24-28): 16 bytes with padding.
type CodeTaxid struct {
	Code  uint64
	Taxid uint32
	_     uint32
}

// SortCodeTaxids replaces sorts.Quicksort(CodeTaxidSlice(mt)): sort.go:268,331,457; split.go:305.
func (x *Ctx) SortCodeTaxids(mt []CodeTaxid, keyBits int) error {
	if len(mt) < 2 {
		return nil
	}
	return x.err(C.ukm_sort_codetaxid16(x.c, unsafe.Pointer(&mt[0]), C.size_t(len(mt)), C.int(keyBits)))
}

func (x *Ctx) nway(call func(in *C.ukm_span, out *C.ukm_span) C.int, sets []Set, outCap int, withTaxid bool) ([]uint64, []uint32, error) {
	in, free := spans(sets)
	defer free()
	codes := make([]uint64, outCap)
	var taxids []uint32
	if withTaxid {
		taxids = make([]uint32, outCap)
	}
	out := outSpan(codes, taxids)
	st := call(in, &out)
	runtime.KeepAlive(sets)
	if e := x.err(st); e != nil {
		return nil, nil, e
	}
	n := int(out.n)
	if withTaxid {
		return codes[:n], taxids[:n], nil
	}
	return codes[:n], nil, nil
}

func total(sets []Set) int {
	n := 0
	for i := range sets {
		n += len(sets[i].Codes)
	}
	return n
}

// Union replaces union.go:186-208 + the ordered emit 260-305 (`-s`).
func (x *Ctx) Union(sets []Set, flags uint) ([]uint64, []uint32, error) {
	return x.nway(func(in, out *C.ukm_span) C.int {
		return C.ukm_union(x.c, in, C.int(len(sets)), C.uint(flags), out)
	}, sets, total(sets), flags&FTaxid != 0)
}

// Inter replaces inter.go:188-286.
func (x *Ctx) Inter(sets []Set, flags uint) ([]uint64, []uint32, error) {
	return x.nway(func(in, out *C.ukm_span) C.int {
		return C.ukm_inter(x.c, in, C.int(len(sets)), C.uint(flags), out)
	}, sets, len(sets[0].Codes), flags&(FTaxid|FMixTaxid) != 0)
}

// Diff replaces diff.go:136-146, 341-515 and the `-s` emit 566-594.
func (x *Ctx) Diff(sets []Set, flags uint) ([]uint64, []uint32, error) {
	return x.nway(func(in, out *C.ukm_span) C.int {
		return C.ukm_diff(x.c, in, C.int(len(sets)), C.uint(flags), out)
	}, sets, len(sets[0].Codes), flags&FTaxid != 0)
}

// PinnedUint64s returns a []uint64 of n elements in page-locked host memory (ukm_alloc_pinned): host spans in pinned
// memory are what lets SetopsStream (and Inter / Diff / Union on host sets) overlap uploads, kernels and downloads.
// The readers fill it in place (reader.ReadCodeWithTaxid in a batch loop); release it with FreePinned.
func PinnedUint64s(n int) []uint64 {
	p := C.ukm_alloc_pinned(C.size_t(n) * 8)
	if p == nil {
		return nil
	}
	return unsafe.Slice((*uint64)(p), n)
}

// FreePinned releases a slice obtained from PinnedUint64s.
func FreePinned(s []uint64) {
	if cap(s) > 0 {
		C.ukm_free_pinned(unsafe.Pointer(&s[:1][0]))
	}
}

// SetopsStream runs any subset of inter / diff / union (ops: OpInter, OpDiff, OpUnion) over the SAME sets in one call:
// what `unikmer inter`, `unikmer diff` and `unikmer union` over one file list compute (inter.go:188-286,
// diff.go:136-146 + 380-435, union.go:186-208 + 260-305), with every input byte uploaded once -- the library streams key
// ranges so that uploads, kernels and downloads overlap.  Keys only.  outs[k] receives the result of ops[k] and must have
// the capacity of sets[0] (inter, diff) or of all sets together (union); the returned slices are outs[k][:n].
func (x *Ctx) SetopsStream(sets []Set, ops []int, flags uint, outs [][]uint64) ([][]uint64, error) {
	in, free := spans(sets)
	defer free()
	cops := make([]C.int, len(ops))
	for i, o := range ops {
		cops[i] = C.int(o)
	}
	n := len(ops)
	po := (*C.ukm_span)(C.malloc(C.size_t(n) * C.size_t(unsafe.Sizeof(C.ukm_span{}))))
	defer C.free(unsafe.Pointer(po))
	osp := unsafe.Slice(po, n)
	for i := range ops {
		osp[i] = outSpan(outs[i], nil)
	}
	st := C.ukm_setops_stream(x.c, in, C.int(len(sets)), &cops[0], C.int(n), C.uint(flags), po)
	runtime.KeepAlive(sets)
	runtime.KeepAlive(outs)
	if e := x.err(st); e != nil {
		return nil, e
	}
	res := make([][]uint64, n)
	for i := range ops {
		res[i] = outs[i][:int(osp[i].n)]
	}
	return res, nil
}

// Common replaces common.go:220-283, 329-354; threshold as computed at common.go:93-105.
func (x *Ctx) Common(sets []Set, flags uint, threshold uint16) ([]uint64, []uint32, error) {
	return x.nway(func(in, out *C.ukm_span) C.int {
		return C.ukm_common(x.c, in, C.int(len(sets)), C.uint(flags), C.uint16_t(threshold), out)
	}, sets, total(sets), flags&FTaxid != 0)
}

// MergeSorted replaces mergeChunksFile (util-sort.go:227-606).
func (x *Ctx) MergeSorted(mode int, sets []Set, flags uint) ([]uint64, []uint32, error) {
	return x.nway(func(in, out *C.ukm_span) C.int {
		return C.ukm_merge_sorted(x.c, C.int(mode), in, C.int(len(sets)), C.uint(flags), out)
	}, sets, total(sets)+2, flags&FTaxid != 0)
}

// FoldSorted replaces the scan loops sort.go:482-573 / dumpCodes*2File (util-sort.go:35-190).
func (x *Ctx) FoldSorted(mode int, s Set, flags uint) ([]uint64, []uint32, error) {
	return x.nway(func(in, out *C.ukm_span) C.int {
		return C.ukm_fold_sorted(x.c, C.int(mode), in, C.uint(flags), out)
	}, []Set{s}, len(s.Codes)+2, flags&FTaxid != 0)
}

// CountMinimizer replaces sketches.NewMinimizerSketch / NextMinimizer + dedup + sort (count.go:316-317, 358-359, 434-436, 581).
func (x *Ctx) CountMinimizer(bases []byte, recOff []uint64, k, w int, flags uint, maxHash uint64) ([]uint64, error) {
	codes := make([]uint64, len(bases)+1)
	out := outSpan(codes, nil)
	var b *C.uint8_t
	if len(bases) > 0 {
		b = (*C.uint8_t)(unsafe.Pointer(&bases[0]))
	}
	st := C.ukm_count_minimizer(x.c, b, (*C.uint64_t)(unsafe.Pointer(&recOff[0])), C.size_t(len(recOff)-1), C.int(k), C.int(w),
		C.uint(flags), C.uint64_t(maxHash), C.UKM_HOST, &out)
	if e := x.err(st); e != nil {
		return nil, e
	}
	return codes[:int(out.n)], nil
}

// CountSeq replaces the iterator + dedup map + sort of count.go:314-322, 355-437, 531-595.
// bases holds the records back to back (fastx strips line breaks); recOff[r]..recOff[r+1] is record r.
func (x *Ctx) CountSeq(bases []byte, recOff []uint64, k int, flags uint, maxHash uint64) ([]uint64, error) {
	codes := make([]uint64, len(bases)+1)
	out := outSpan(codes, nil)
	var b *C.uint8_t
	if len(bases) > 0 {
		b = (*C.uint8_t)(unsafe.Pointer(&bases[0]))
	}
	st := C.ukm_count_seq(x.c, b, (*C.uint64_t)(unsafe.Pointer(&recOff[0])), C.size_t(len(recOff)-1), C.int(k),
		C.uint(flags), C.uint64_t(maxHash), C.UKM_HOST, &out)
	if e := x.err(st); e != nil {
		return nil, e
	}
	return codes[:int(out.n)], nil
}
