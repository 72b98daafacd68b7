import sys, time
sys.path.insert(0, "/root/repo")
import wann_b200
from wann_b200 import weights as W
w = W.synthetic("trained_like", 1)
for mode in ("bf16", "fp16c", "fp16c"):
    t0 = time.time()
    ev = wann_b200.PolicyEvaluator(w, mode=mode)
    t1 = time.time()
    if mode == "fp16c":
        ev.set_option("gptq", 0); t2 = time.time(); ev.set_option("gptq", 1); t3 = time.time()
        print(mode, "create %.2f s, recalibrate without gptq %.2f s, with %.2f s" % (t1 - t0, t2 - t1, t3 - t2), flush=True)
    else:
        print(mode, "create %.2f s" % (t1 - t0), flush=True)
    ev.close()
