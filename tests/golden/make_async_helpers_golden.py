#!/usr/bin/env python
"""Generate tests/golden/reference_async_helpers.npz with the REFERENCE's own pure-torch helpers of
src/dagr/asynchronous (build container only; third-party imports replaced by tests/golden/ref_shim.py)."""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
from tests.golden import ref_shim                                        # noqa: E402

ref_shim.install()
import dagr.asynchronous.cartesian as ref_cart                            # noqa: E402
import dagr.asynchronous.max_pool as ref_pool                             # noqa: E402
import dagr.asynchronous.base.utils as ref_utils                          # noqa: E402

g = torch.Generator().manual_seed(5)
n, E = 600, 4000
W, H = 320, 215
pos = torch.stack([torch.randint(0, W, (n,), generator=g) / W, torch.randint(0, H, (n,), generator=g) / H,
                   torch.sort(torch.rand(n, generator=g)).values], 1).float()
ei = torch.randint(0, n, (2, E), generator=g)
out = dict(pos=pos.numpy(), edge_index=ei.numpy().astype(np.int32))
edge_attr = getattr(ref_cart, "__edge_attr")
for tag, mx in (("a", 2 * float(int(0.01 * W + 2) / W)), ("b", 0.1)):
    out[f"edge_attr_{tag}"] = edge_attr(pos, ei, True, mx).numpy()
    out[f"max_{tag}"] = np.array(mx)


class _M:                                                                 # what the helpers read from the pooling module
    voxel_size = torch.tensor([1.0 / 28, 1.0 / 20, 1.0, 1.0])
    dim = 2


class _T:
    max = 0.1


cluster = getattr(ref_pool, "__get_global_cluster_index")(_M, pos[:, :2])
out["voxel_size"] = _M.voxel_size.numpy()
out["global_cluster"] = cluster.numpy().astype(np.int32)
uniq, cl = torch.unique(cluster, sorted=True, return_inverse=True)
out["cluster"] = cl.numpy().astype(np.int32)
pe = ref_pool.pool_edge(cl, ei, False)
out["pooled_edges"] = pe.numpy().astype(np.int32)
hom = ref_utils._to_hom(pos)
out["hom"] = hom.numpy()
sums = torch.zeros(len(uniq), 4).index_add_(0, cl, hom)
cpos = ref_utils._from_hom(sums)
out["from_hom"] = cpos.numpy()
out["cpos"] = cpos.numpy()
out["pooled_attr"] = ref_pool.compute_attrs(_T, pe, cpos).numpy()
np.savez_compressed(HERE / "reference_async_helpers.npz", **out)
print("wrote", sorted(out))
