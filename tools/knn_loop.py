import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import seurat_b200 as sb
from seurat_b200.synthetic import gaussian_mixture
X,_ = gaussian_mixture(1_000_000, 50, 1)
ctx = sb.Context(0)
ref = torch.from_numpy(X).cuda()
idx0 = torch.empty((1_000_000, 20), dtype=torch.int32, device="cuda"); dist = torch.empty((1_000_000, 20), dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
NC = int(sys.argv[1]) if len(sys.argv) > 1 else 0
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 8
for r in range(REPS):
    t0=time.perf_counter(); st = ctx.dev_knn(ref, 1_000_000, None, 1_000_000, 50, 20, idx0, dist, 0, NC); dt=time.perf_counter()-t0
    print(f"{dt*1e3:.1f} ms cand {st['ms_knn_cand']:.1f} rerank {st['ms_knn_rerank']:.1f} fallback {st['ms_knn_fallback']:.1f} uncert {st['n_uncertified']} tiles {st['knn_tiles_visited']}")
# sanity: index range and a sampled brute-force comparison (float64 torch)
torch.cuda.synchronize()
bad = ((idx0 < 0) | (idx0 >= 1_000_000)).nonzero()
print("out-of-range entries:", bad.shape[0], bad[:5].tolist(), idx0[bad[:5, 0]].tolist() if bad.shape[0] else "")
rows = torch.randint(0, 1_000_000, (256,), device="cuda")
if bad.shape[0]:
    rows[: min(8, bad.shape[0])] = bad[:8, 0]
d2 = torch.cdist(ref[rows], ref, compute_mode="donot_use_mm_for_euclid_dist") ** 2
want = d2.topk(20, largest=False).indices
got = idx0[rows].long()
same = (want.sort(1).values == got.sort(1).values).all(1)
print("sampled rows with the same neighbour set:", int(same.sum()), "of", rows.numel())
