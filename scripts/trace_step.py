"""In-situ kernel timeline of one cfg2 prefill step (CUDA-graph replay) via torch.profiler/CUPTI: per-kernel totals plus the
start/duration/gap sequence of one encoder layer and one Llama layer.  Diagnostic only - not a bench number."""
import json, os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from ultravox_b200.config import preset
from ultravox_b200.engine import PrefillEngine
from ultravox_b200.model import UltravoxModel

cfg = preset("v0_5_8b")
dev = torch.device("cuda", 0)
model = UltravoxModel(cfg, device=dev).init_random_(seed=42)
wl = bench.workload(cfg, 30.0)
n = wl["n"]
eng = PrefillEngine(model, wl["n"], wl["input_ids"], wl["start"], wl["tok_len"], wl["abs"])
w = torch.from_numpy(np.random.default_rng(1000).standard_normal(n).astype(np.float32))[None]
w = torch.nn.functional.pad(w, (0, eng.L - w.shape[1])).to(dev)
for _ in range(3):
    eng.wave.copy_(w); eng.run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        eng.wave.copy_(w); eng.run()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower() and "memset" not in e.name.lower()]
evs.sort(key=lambda e: e.time_range.start)
half = len(evs) // 2
step = evs[half:]
t0 = step[0].time_range.start
agg = collections.OrderedDict()
for e in step:
    k = e.name.split("(")[0][-60:]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += e.time_range.end - e.time_range.start
span = step[-1].time_range.end - t0
busy = sum(v[1] for v in agg.values())
print(f"kernels {len(step)} span_us {span:.1f} busy_us {busy:.1f} gaps_us {span - busy:.1f}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} {v[0]:4d} {v[1]:9.1f} {v[1] / v[0]:8.2f} {100 * v[1] / span:5.1f}%")
# Under programmatic dependent launch a kernel STARTS (launch, set-up, first weight boxes) while its predecessor still runs, so
# start-to-end durations overlap and over-count.  What a kernel adds to the step is the time from its predecessor's END to its
# own END: those deltas sum to the span exactly.
adv = collections.OrderedDict()
prev_end = t0
for e in step:
    k = e.name.split("(")[0][-60:]
    end = max(e.time_range.end, prev_end)
    a = adv.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += end - prev_end
    prev_end = end
print("-- end-to-end advance per kernel (sums to the span)")
for k, v in sorted(adv.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} {v[0]:4d} {v[1]:9.1f} {v[1] / v[0]:8.2f} {100 * v[1] / span:5.1f}%")
def dump(lo, hi):
    prev = step[lo - 1].time_range.end if lo > 0 else None
    for e in step[lo:hi]:
        s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
        gap = (e.time_range.start - prev) if prev is not None else 0.0
        adv_ = (e.time_range.end - prev) if prev is not None else d
        print(f"  +{s:9.1f}us gap {gap:5.1f} dur {d:7.2f} advance {adv_:7.2f}  {e.name.split('(')[0][-50:]}")
        prev = max(e.time_range.end, prev) if prev is not None else e.time_range.end
print("-- encoder layer 10"); 
names = [e.name for e in step]
ln = [i for i, nme in enumerate(names) if "layernorm" in nme]
dump(ln[20], ln[22] + 1)
an = [i for i, nme in enumerate(names) if "attn_llm_tc" in nme or "attn_fwd_kernel<128>" in nme]
print("-- llama layer 10")
if len(an) > 11:
    per = an[11] - an[10]
    dump(an[10] - 3, an[10] - 3 + per)
