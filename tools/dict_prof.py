"""Times the pieces of the multi-GPU barcode dictionary union (cratac/multi.py:global_columns) on ONE GPU with
synthetic per-rank tables (tools only)."""
import sys, time
sys.path.insert(0, 'cellranger-atac_b200')
import numpy as np, torch
from cratac import _ffi, multi

world, nb = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 210000
ctx = _ffi.Context([("chr1", 1000)])
dev = torch.device("cuda")
g = torch.Generator(device=dev); g.manual_seed(1)
universe = torch.randint(0, 1 << 40, (nb,), generator=g, device=dev, dtype=torch.int64)
hashes = torch.randint(-(1 << 62), 1 << 62, (nb,), generator=g, device=dev, dtype=torch.int64)
keys_all = torch.zeros((world, nb), dtype=torch.int64, device=dev)
first_all = torch.zeros_like(keys_all); hash_all = torch.zeros_like(keys_all)
valid = torch.zeros((world, nb), dtype=torch.bool, device=dev)
for r in range(world):
    perm = torch.randperm(nb, generator=g, device=dev)[: nb - 1000 * r]
    n = perm.numel()
    keys_all[r, :n] = universe[perm]; hash_all[r, :n] = hashes[perm]
    first_all[r, :n] = torch.arange(n, device=dev) + r * 40_000_000
    valid[r, :n] = True

def py2(h):
    out = torch.empty(h.numel(), dtype=torch.int32, device=dev)
    torch.cuda.current_stream().synchronize()
    ctx.py2_set_order_hashes_device(h.numel(), h.data_ptr(), out.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    return out

def timed(fn, *a):
    torch.cuda.synchronize(); t = time.time(); r = fn(*a); torch.cuda.synchronize(); return r, 1e3 * (time.time() - t)

for it in range(3):
    _, t_all = timed(multi.global_columns, torch, keys_all, first_all, hash_all, valid, py2)
    fv = valid.reshape(-1)
    (keys), t1 = timed(lambda: keys_all.reshape(-1)[fv])
    (u, inv), t2 = timed(lambda: torch.unique(keys, return_inverse=True))
    h = hashes.clone()
    _, t3 = timed(py2, h)
    _, t4 = timed(lambda: torch.argsort(first_all[0]))
    print("world %d: global_columns %.2f ms | mask-index %.2f | unique %.2f | py2 device %.2f | argsort %.2f" % (world, t_all, t1, t2, t3, t4))
