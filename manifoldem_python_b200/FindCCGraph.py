"""The S2 pairwise-distance step of the CC-graph construction (modules/FindCCGraph.py:227-273, called at :296 with the
thresholded bin centres S20_th) — SURVEY.md §8f rank 4.  Only this function of FindCCGraph is on the
"orientations -> embeddings" path; the graph bookkeeping around it (CreateGraphStruct, pruning, belief propagation)
is out of scope.  A maintainer binds it with  `FindCCGraph.CalcPairwiseDistS2 = manifoldem_python_b200.FindCCGraph.
CalcPairwiseDistS2`  (INTEGRATION.md).

    CalcPairwiseDistS2(X)                    X: 3 x N coordinates -> (pwDotProd (N,N), pwDist (N,N))
    CalcPairwiseDistS2(X, UIdxs, VIdxs)      the two index sets of :253-262

Device: `mem_s2_pairwise_host` (s2.cu k_s2_pairwise, float64, one thread per pair)."""
import numpy as np

from . import _lib
from .getDistanceCTF_local_Conj9combinedS2 import _ctx


def CalcPairwiseDistS2(X, *argv, ctx=None):
    if not (len(argv) == 0 or len(argv) == 2):
        raise AssertionError('wrong nmber of arguments')                      # :244-251
    X = np.asarray(X)
    if not argv:
        UIdxs = VIdxs = np.arange(X.shape[1])
    else:
        UIdxs, VIdxs = argv
    U = np.ascontiguousarray(np.atleast_2d(X[:, UIdxs].T), dtype=np.float64)   # one point per row
    V = np.ascontiguousarray(np.atleast_2d(X[:, VIdxs].T), dtype=np.float64)
    if U.shape[1] != 3 or V.shape[1] != 3:
        raise ValueError('CalcPairwiseDistS2 expects a 3 x N coordinate matrix')
    if U.shape[0] != V.shape[0]:
        # np.sum(U*U,0).T + np.sum(V*V,0) (:267) is an elementwise sum of two 1-D arrays: NumPy raises here too
        raise ValueError('operands could not be broadcast together with shapes (%d,) (%d,)' % (U.shape[0], V.shape[0]))
    lib = _lib.load()
    ctx = ctx or _ctx()
    dot = np.empty((U.shape[0], V.shape[0]))
    dist = np.empty((U.shape[0], V.shape[0]))
    _lib.check(lib.mem_s2_pairwise_host(ctx.handle, U.ctypes.data, U.shape[0], V.ctypes.data, V.shape[0],
                                        dot.ctypes.data, dist.ctypes.data))
    return dot, dist
