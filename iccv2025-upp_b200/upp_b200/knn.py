"""Mirror of knn_cuda.KNN (KNN_CUDA 0.2; reference models/Point_MAE_unify.py:16,56,69)."""
import torch
import torch.nn as nn

from . import ops


class KNN(nn.Module):
    """KNN(k, transpose_mode=False)(ref, query) -> (D, I).

    transpose_mode=True : ref (B,N,3), query (B,Q,3) -> D, I (B,Q,k)   (what UPP uses)
    transpose_mode=False: ref (B,3,N), query (B,3,Q) -> D, I (B,k,Q)
    D: Euclidean distances, ascending; I: int64, 0-based, equal distances keep the lower index
    first.  Runs under no_grad like upstream; inputs are cast with .float().  One kernel launch
    for the whole batch (upstream: a Python loop over clouds).  Only 3-D points are supported.
    """

    def __init__(self, k, transpose_mode=False):
        super().__init__()
        self.k = k
        self._t = transpose_mode

    def forward(self, ref, query):
        assert ref.size(0) == query.size(0), "ref.shape={} != query.shape={}".format(ref.shape, query.shape)
        with torch.no_grad():
            if self._t:
                r, q = ref.float().contiguous(), query.float().contiguous()
            else:
                r, q = ref.float().transpose(1, 2).contiguous(), query.float().transpose(1, 2).contiguous()
            if r.size(2) != 3 or q.size(2) != 3:
                raise NotImplementedError("upp_b200.KNN supports 3-D points only (the UPP Group divider)")
            D, I = ops.knn(r, q, self.k)
            if not self._t:
                D, I = D.transpose(1, 2).contiguous(), I.transpose(1, 2).contiguous()
        return D, I
