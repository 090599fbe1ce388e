"""N-way inter / diff with and without per-tile work (UKM_NWAY_NULL=1: no work, =2: load pipeline only; results not valid).
Needs a measurement build of the library:  make -C unikmer_b200/csrc clean all EXTRA=-DUKM_MEASURE"""
import os, sys, json, torch
sys.path.insert(0, os.getcwd())
from tools.exp_nway import timed
from unikmer_b200 import Engine
eng = Engine(0); stream = torch.cuda.Stream(); eng.use_stream(stream.cuda_stream)
with torch.cuda.stream(stream):
    U = 10**9
    files = [eng.synth_member_file(0, U, U, 3, 4, f).clone() for f in range(8)]
    out = torch.empty(int(files[0].shape[0]) + 16, dtype=torch.int64, device="cuda")
    os.environ["UKM_NWAY_FILTER"] = "1"
    for null in ("2", "1"):
        os.environ["UKM_NWAY_NULL"] = null
        for cfg in ("0", "4", "2"):
            os.environ["UKM_NWAY_CFG"] = cfg
            for name, fn in (("inter", eng.inter), ("diff", eng.diff)):
                eng.stats_reset(); eng.stats_enable(True)
                ms = timed(stream, lambda: fn(files, out=out), reps=3)
                eng.stats_enable(False)
                st = eng.stats()
                print(json.dumps({"null": null, "cfg": cfg, "op": name, "ms": round(ms, 3), "n_out": int(fn(files, out=out)[0].shape[0]),
                                  "k": {k: round(v["ms"] / max(v["launches"], 1), 3) for k, v in st.items()}}), flush=True)
    # the union through the same shapes
    outu = torch.empty(min(sum(int(f.shape[0]) for f in files), U) + 16, dtype=torch.int64, device="cuda")
    os.environ["UKM_NWAY_NULL"] = "0"
    for cfg in ("0", "5", "6"):
        os.environ["UKM_NWAY_CFG"] = cfg
        ms = timed(stream, lambda: eng.union(files, out=outu), reps=3)
        print(json.dumps({"op": "union", "cfg": cfg, "ms": round(ms, 3)}), flush=True)
