"""Per-tile timeline of the fp64 grouped GEMM on the bench workload (t10_ggemm_set_trace).
Prints: kernel span, per-SM busy fraction, when CTAs/SMs go idle, tile duration against the cost model.
    TOR10_GGEMM_SCHED=slots|smbins|dynamic python tools/trace_ggemm.py [chi]
"""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import tor10_b200 as T
from tor10_b200 import _plan, _lib
import bench
from helpers import psym

chi = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
A, B = bench.c3a_specs(chi)
a = psym(T, A, seed=1003, device=dev); b = psym(T, B, seed=1004, device=dev)
plan = _plan.get_contract_plan(a, b)
g = plan.gemm
for _ in range(3):
    T.Contract(a, b)
torch.cuda.synchronize()
sk = getattr(g, "streamk", False) or getattr(g, "tma", False)
ntl = g.sk_npieces if sk else g.ntiles
trace = torch.zeros((ntl, 4), dtype=torch.int64, device=dev)
lib = _lib.load()
lib.t10_ggemm_set_trace(trace.data_ptr())
T.Contract(a, b)
torch.cuda.synchronize()
lib.t10_ggemm_set_trace(None)
tr = trace.cpu().numpy()
tiles = (g.d_sk_tiles if sk else g.d_tiles).cpu().numpy()[:ntl]
pr = g.problems
sm, cta, t0, t1 = tr[:, 0], tr[:, 1], tr[:, 2], tr[:, 3]
ok = t1 > 0
assert ok.all(), "untraced tiles: %d" % (~ok).sum()
start = t0.min(); span = (t1.max() - start) * 1e-3
dur = (t1 - t0) * 1e-3
kk = pr["k"][tiles[:, 0]].astype(np.int64)
mm = np.minimum(64, pr["m"][tiles[:, 0]] - tiles[:, 1]); nn = np.minimum(64, pr["n"][tiles[:, 0]] - tiles[:, 2])
gran = 32 if g.tile_cfg in (40, 42, 46) else 8
cost = (-(-kk // 16)) * 4 * (-(-mm // gran)) * (-(-nn // gran)) * (gran // 8) ** 2       # DMMAs per warp-set
if g.tile_cfg == 46:      # mixed shapes: every tile is four warp tiles wide; cost = k-steps x 64 fragments x 4
    cost = (-(-kk // 16)) * 4 * 64
if sk:
    pcs = g.d_sk_pieces.cpu().numpy()
    cost = (pcs[:, 1] - pcs[:, 0]).astype(np.int64) * 4 * 64
print("sched=%s cfg=%d tiles=%d span=%.1f us  sum(tile dur)=%.0f us  CTAs seen=%d SMs seen=%d"
      % (("stream-K" if sk else os.environ.get("TOR10_GGEMM_SCHED", "slots")), g.tile_cfg, ntl, span, dur.sum(), len(set(cta)), len(set(sm))))
# per-SM: union of busy intervals, last end
sm_end = {}; sm_work = {}
for s_, e_, c_ in zip(sm, t1, cost):
    sm_end[s_] = max(sm_end.get(s_, 0), e_); sm_work[s_] = sm_work.get(s_, 0) + c_
ends = np.array([(sm_end[s_] - start) * 1e-3 for s_ in sorted(sm_end)])
work = np.array([sm_work[s_] for s_ in sorted(sm_end)], dtype=np.float64)
print("SM finish time: min %.1f  p25 %.1f  median %.1f  p75 %.1f  max %.1f us" % (ends.min(), np.percentile(ends, 25), np.median(ends), np.percentile(ends, 75), ends.max()))
print("SM work (DMMA model): min %.0f mean %.0f max %.0f  -> max/mean %.3f" % (work.min(), work.mean(), work.max(), work.max() / work.mean()))
rate = work / ends
print("SM rate (model DMMA per us until its finish): min %.0f median %.0f max %.0f" % (rate.min(), np.median(rate), rate.max()))
# CTA-level
cta_end = {}
for c_, e_ in zip(cta, t1):
    cta_end[c_] = max(cta_end.get(c_, 0), e_)
ce = np.array([(v - start) * 1e-3 for v in cta_end.values()])
print("CTA finish: p10 %.1f median %.1f p90 %.1f max %.1f us" % (np.percentile(ce, 10), np.median(ce), np.percentile(ce, 90), ce.max()))
# start skew
cta_start = {}
for c_, s_ in zip(cta, t0):
    cta_start[c_] = min(cta_start.get(c_, 1 << 62), s_)
cs = np.array([(v - start) * 1e-3 for v in cta_start.values()])
print("CTA first-tile start: median %.1f p90 %.1f max %.1f us" % (np.median(cs), np.percentile(cs, 90), cs.max()))
# tile duration vs cost: us per 1000 DMMA-steps, by phase
full = cost == cost.max()
print("tiles at max cost (%d DMMA): %d, duration mean %.1f min %.1f max %.1f us" % (cost.max(), full.sum(), dur[full].mean(), dur[full].min(), dur[full].max()))
order = np.argsort(t0)
q = len(order) // 4
for i, name in enumerate(("first quarter", "second", "third", "last quarter")):
    idx = order[i * q:(i + 1) * q] if i < 3 else order[3 * q:]
    print("  %-14s started %.1f..%.1f us: us per kDMMA = %.2f" % (name, (t0[idx].min() - start) * 1e-3, (t0[idx].max() - start) * 1e-3, dur[idx].sum() / cost[idx].sum() * 1e3))
# occupancy over time: number of active CTAs sampled every 5 us
for tt in range(0, int(span) + 5, 10):
    ts = start + tt * 1000
    act = ((t0 <= ts) & (t1 > ts)).sum()
    sms = len(set(sm[(t0 <= ts) & (t1 > ts)]))
    print("  t=%3d us: active tiles %3d on %3d SMs" % (tt, act, sms))
np.save(os.path.join(ROOT, "gpurun_out", "trace_%s.npy" % os.environ.get("TOR10_GGEMM_SCHED", "slots")), np.concatenate([tr, tiles, cost[:, None]], axis=1))
