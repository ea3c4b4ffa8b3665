"""z-slab sharding of the FDTD domain across GPUs (one process per GPU).

Mirrors the reference's MPI domain split (openems_fdtd_mpi.cpp:201-299, engine_mpi.cpp:84-210)
for one axis: rank r owns planes [z_r, z_{r+1}) and holds one ghost plane per interface.
After the E half-step the lowest owned plane's tangential E goes DOWN into the lower rank's
ghost-E plane; after the H half-step the highest owned plane's tangential H goes UP.
Host logic only; the transfers themselves are device-to-device writes into NVLink peer
memory issued by the halo kernels of libopenems_b200.so.
"""


def slab_range(nz, world, rank, pml_lo=0, pml_hi=0, pml_weight=0.0):
    """owned plane range [zb, ze) of `rank`.  With pml_weight > 0 the end slabs (whose z-PML
    planes cost more per plane) get proportionally fewer planes."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    if nz < 2 * world:
        raise ValueError("need at least two planes per slab")
    cost = [1.0] * nz
    for k in range(nz):
        if k <= pml_lo and pml_lo:
            cost[k] += pml_weight
        if k >= nz - 1 - pml_hi and pml_hi:
            cost[k] += pml_weight
    total = sum(cost)
    bounds = [0]
    acc, k = 0.0, 0
    for r in range(1, world):
        target = total * r / world
        while k < nz and acc + cost[k] <= target + 1e-9:
            acc += cost[k]
            k += 1
        k = max(k, bounds[-1] + 2)
        k = min(k, nz - 2 * (world - r))
        acc = sum(cost[:k])
        bounds.append(k)
    bounds.append(nz)
    return bounds[rank], bounds[rank + 1]


def held_range(nz, zb, ze):
    """planes held by a slab engine: owned planes plus one ghost plane per interface"""
    return zb - (1 if zb > 0 else 0), ze + (1 if ze < nz else 0)


def link_engines_in_process(engines):
    """neighbour wiring for several slab engines living in ONE process (tests, single host
    process driving several GPUs): direct peer pointers instead of IPC handles"""
    for r, e in enumerate(engines):
        e.LinkPeers(engines[r - 1] if r > 0 else None, engines[r + 1] if r + 1 < len(engines) else None)


def link_engines_distributed(eng, dist, rank, world):
    """neighbour wiring with one process per GPU: all-gather the CUDA IPC handles over the
    torch.distributed process group (any backend), open the two neighbours' handles"""
    blobs = [None] * world
    dist.all_gather_object(blobs, eng.ExportIPC())
    eng.OpenPeers(blobs[rank - 1] if rank > 0 else None, blobs[rank + 1] if rank < world - 1 else None)
    dist.barrier()
