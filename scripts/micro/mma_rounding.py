"""How does tcgen05.mma (kind::f16, fp32 accumulate in TMEM) round?  Probes through the library's own one-layer
diagnostic entry (tdmpc2_debug_layer, raw linear output) with hand-made weights / inputs that are exact in fp16, so
that every deviation from the exact sum is the tensor core's own rounding:

  A  inside ONE MMA (K = 16): 1 + 15 products of 1.5 * 2^-j   -> how many bits below the largest product survive,
     and in which direction the dropped bits are rounded (both signs);
  B  across MMAs: accumulator (+-1) + one later product of +-1.5 * 2^-j  -> rounding of the accumulate step;
  C  many small same-sign products added to a large accumulator over 64 K-steps -> drift per step.

    python scripts/micro/mma_rounding.py        (GPU box)
"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from tdmpc2_b200.config import workload
from tdmpc2_b200.synth import synth_state_dict
from tdmpc2_b200.planner import Planner

cfg = workload("tiny", num_envs=1)
sd = synth_state_dict(cfg, seed=1)
K = N = 64
LI = 3            # layer index of _dynamics.1 in the packed table: enc.0, enc.1, dyn.0, dyn.1


def run(W, X, engine="tcgen05"):
    sd2 = dict(sd)
    sd2["_dynamics.1.weight"] = W.float()
    sd2["_dynamics.1.bias"] = torch.zeros(N)
    pl = Planner(cfg, 1, "cuda:0", engine=engine)
    pl.pack(sd2)
    y = pl.debug_layer(LI, 0, X.float().cuda(), N).cpu().double()
    del pl
    return y


def ulps(y, ref):            # deviation in units of the fp32 ulp of the exact value
    u = torch.tensor([2.0 ** (math.floor(math.log2(abs(r))) - 23) if r != 0 else 1.0 for r in ref.tolist()], dtype=torch.float64)
    return (y - ref) / u


for eng in ("tcgen05", "simt"):
    print(f"===== engine {eng}")
    # ---- A: one MMA (k = 0..15), column n <-> exponent j = n + 1
    for sign in (+1, -1):
        W = torch.zeros(N, K, dtype=torch.float64)
        W[:, 0] = 1.0
        for n in range(40):
            W[n, 1:16] = sign * 1.5 * 2.0 ** -(n + 1)
        X = torch.zeros(1, K, dtype=torch.float64); X[0, :16] = 1.0
        y = run(W, X, eng)[0, :40]
        ref = 1.0 + sign * 15 * 1.5 * 2.0 ** -torch.arange(1, 41, dtype=torch.float64)
        ref32 = ref.float().double()
        print(f"A sign={sign:+d}: j : (y - exact)/ulp  [RN would give (ref32-exact)/ulp]")
        print("   " + "  ".join(f"{j + 1}:{d:+.2f}[{r:+.2f}]" for j, (d, r) in enumerate(zip(ulps(y, ref).tolist(), ulps(ref32, ref).tolist())) if j >= 14 and j < 34))
    # ---- B: accumulator from the first K-step (+-1), one product in the second K-step
    for s0 in (+1, -1):
        for s1 in (+1, -1):
            W = torch.zeros(N, K, dtype=torch.float64)
            W[:, 0] = s0 * 1.0
            for n in range(40):
                W[n, 16] = s1 * 1.5 * 2.0 ** -(n + 1)
            X = torch.zeros(1, K, dtype=torch.float64); X[0, 0] = 1.0; X[0, 16] = 1.0
            y = run(W, X, eng)[0, :40]
            ref = s0 * 1.0 + s1 * 1.5 * 2.0 ** -torch.arange(1, 41, dtype=torch.float64)
            print(f"B acc={s0:+d} add={s1:+d}*1.5*2^-j: " + "  ".join(f"{j + 1}:{d:+.2f}" for j, d in enumerate(ulps(y, ref).tolist()) if 20 <= j < 28))
    # ---- C: drift: accumulator 1.0 then 48 later K=16 steps... (K = 64 gives 3 more steps); each adds 16 x 2^-j * 1.25
    for sign in (+1, -1):
        W = torch.zeros(N, K, dtype=torch.float64)
        W[:, 0] = 1.0
        for n in range(40):
            W[n, 16:64] = sign * 1.25 * 2.0 ** -(n + 1)
        X = torch.ones(1, K, dtype=torch.float64); X[0, 1:16] = 0.0
        y = run(W, X, eng)[0, :40]
        ref = 1.0 + sign * 48 * 1.25 * 2.0 ** -torch.arange(1, 41, dtype=torch.float64)
        print(f"C sign={sign:+d} (1 + 48 products of 1.25*2^-j over 3 K-steps): " + "  ".join(f"{j + 1}:{d:+.2f}" for j, d in enumerate(ulps(y, ref).tolist()) if 16 <= j < 32))
