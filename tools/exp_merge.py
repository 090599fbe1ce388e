import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unikmer_b200 import Engine
from tools.microbench import timed
eng = Engine(0); stream = torch.cuda.Stream(); eng.use_stream(stream.cuda_stream)
with torch.cuda.stream(stream):
    U = 10**9
    a = eng.synth_member_file(0, U, U, 3, 4, 0).clone(); b = eng.synth_member_file(0, U, U, 3, 4, 1).clone()
    for pipe in ("0", "2", "off"):
        os.environ["UKM_SETOP_PIPE"] = pipe; os.environ["UKM_SETOP_SKEW"] = "0"
        r = {}
        for name, fn in (("merge", lambda: eng.merge([a, b])), ("union", lambda: eng.union([a, b])), ("inter", lambda: eng.inter([a, b])), ("diff", lambda: eng.diff([a, b]))):
            eng.stats_reset(); eng.stats_enable(True)
            ms = timed(stream, fn, reps=3)
            eng.stats_enable(False)
            st = eng.stats()
            r[name] = {k: round(v["ms"] / v["launches"], 3) for k, v in st.items()}
        print(json.dumps({"pipe": pipe, **r}), flush=True)
