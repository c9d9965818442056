"""pytorch3d.ops.knn_points, the one call of the reference (gaussian_model.py:170: K = 4, self included): exact brute
force, distances as sums of squared differences (what pytorch3d's CUDA kernel computes), ascending, chunked."""
import collections

import torch

_KNN = collections.namedtuple("KNN", "dists idx knn")


def knn_points(p1, p2, K=1, chunk=1024, **kw):
    assert p1.dim() == 3 and p1.shape[0] == 1 and p2.shape[0] == 1, "shim covers the reference's (1, N, 3) call only"
    a, b = p1[0], p2[0]
    N = a.shape[0]
    dists = torch.empty(N, K, dtype=a.dtype, device=a.device)
    idx = torch.empty(N, K, dtype=torch.long, device=a.device)
    bc = (b - b.mean(0, keepdim=True)).double()
    bs = (bc * bc).sum(-1)
    kk = min(2 * K + 2, b.shape[0])
    for s in range(0, N, chunk):
        ac = (a[s:s + chunk] - b.mean(0, keepdim=True)).double()
        d2 = (ac * ac).sum(-1)[:, None] + bs[None, :] - 2.0 * ac @ bc.T     # ranking only
        cand = torch.topk(d2, kk, dim=1, largest=False).indices
        diff = a[s:s + chunk, None, :] - b[cand]
        exact = (diff * diff).sum(-1)
        v, j = torch.topk(exact, K, dim=1, largest=False)
        dists[s:s + chunk] = v
        idx[s:s + chunk] = torch.gather(cand, 1, j)
    return _KNN(dists[None], idx[None], None)
