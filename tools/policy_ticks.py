"""Per-phase clock64 ticks of CTA 0 of the fused policy kernel (library built with -DUAVCA_POLICY_TIMING, loaded via
UAVCA_LIB): per iteration sample of the previous tile + wait for the tensor pipe, layer-1 epilogue, barrier, heads of the
previous tile, next layer-1 operand + barrier; total; tiles of CTA 0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_uav_collision_avoidance_b200 as G

M = int(sys.argv[1]) if len(sys.argv) > 1 else 163840
torch.manual_seed(0)
p = G.GaussianPolicy(10, 2).cuda()
f = G.FusedGaussianPolicy(p)
obs = torch.rand(M, 10, device="cuda") * 2 - 1
out = torch.empty(M, 2, device="cuda")
head = torch.zeros(M, 4, device="cuda")
for _ in range(3):
    f.act(obs, out=out, head=head)
torch.cuda.synchronize()
names = ["sample+sync2+mma_wait", "epi1.fence", "sync1", "heads", "sync_epi", "tail", "epi1.ldtm", "epi1.cvt+sts", "-", "-"]
mnames = ["L1_issue+mma_wait", "stage.loads", "sync1", "mma2_issue", "sync2", "tail", "stage.cvt+sts", "stage.fence", "-", "-"]
for who, off, nm in (("epilogue warp 0 (SMSP 0)", 0, names), ("epilogue warp 1 (SMSP 1)", 12, names), ("MMA warp", 24, mnames)):
    h = head.flatten()[off:off + 12].cpu().tolist()
    t = max(h[11], 1.0)
    print(who, "per tile:", {n: round(v / t) for n, v in zip(nm[:8], h[:8])}, "total", int(h[10]), "tiles", int(h[11]))
