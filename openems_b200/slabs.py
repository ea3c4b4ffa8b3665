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


# ---------------------------------------------------------------------------------------------
# readout on z-slab engines: every slab evaluates the part of a dump box / mode plane / probe set it
# owns (template: the reference's MPI ranks each process their sub-domain, openems_fdtd_mpi.cpp:201-299);
# the pieces are put together here.  `engines` = the slab engines of ONE process in z order; the
# *_distributed variants take this rank's engine and a torch.distributed-like module.
# ---------------------------------------------------------------------------------------------
def exchange_ghosts(engines):
    """complete every slab's ghost planes for an interpolating readout (enqueue only, all slabs first: a single host
    thread must not block on one slab before the others have been enqueued)"""
    for e in engines:
        e.ExchangeGhosts()


def read_dump_slabs(engines, dump_ids):
    """ProcessFields::CalcField over slabs: dump_ids[r] = the id AddDump returned on engines[r] (same box on all);
    returns the whole box {3, nz, ny, nx}"""
    import numpy as np
    exchange_ghosts(engines)
    return np.concatenate([e.ReadDump(d) for e, d in zip(engines, dump_ids)], axis=1)


def accumulate_fd_slabs(engines, fd_ids, weights):
    exchange_ghosts(engines)
    for e, f in zip(engines, fd_ids):
        e.AccumulateFD(f, weights)


def read_fd_slabs(engines, fd_ids):
    import numpy as np
    parts = [e.ReadFD(f) for e, f in zip(engines, fd_ids)]
    return np.concatenate([p[0] for p in parts], axis=2), parts[0][1]


def combine_mode_match(parts):
    """parts = (value, _, purity) of every slab -> (value, value^2/purity)"""
    value = sum(p[0] for p in parts)
    purity = sum(p[2] for p in parts)
    return value, (value * value / purity if purity != 0.0 else 0.0)


def mode_match_slabs(engines, mode_ids):
    exchange_ghosts(engines)
    return combine_mode_match([e.ReadModeMatchRaw(m) for e, m in zip(engines, mode_ids)])


def combine_steadystate(parts, period):
    """parts = SteadyStateRaw() of every slab -> (last_diff, checks): energies add up, records are concatenated"""
    import ctypes as C
    import numpy as np
    from ._lib import load_library as load
    L = load()
    info = parts[0][0].copy()
    en = np.sum([p[1] for p in parts], axis=0)
    snap = np.ascontiguousarray(np.concatenate([p[2] for p in parts], axis=1), np.float64)
    d = C.c_double()
    dp, up = C.POINTER(C.c_double), C.POINTER(C.c_uint)
    rc = L.oems_cuda_steadystate_eval(int(period), snap.shape[1], info.ctypes.data_as(up), en.ctypes.data_as(dp),
                                      snap.ctypes.data_as(dp) if snap.size else None, C.byref(d))
    if rc:
        raise RuntimeError("oems_cuda_steadystate_eval failed")
    return d.value, int(info[0])


def steadystate_slabs(engines, period):
    return combine_steadystate([e.SteadyStateRaw() for e in engines], period)


def read_dump_distributed(eng, dump_id, dist, world):
    """one process per GPU: every rank reads its piece, rank pieces are all-gathered and concatenated along z"""
    import numpy as np
    mine = eng.ReadDump(dump_id)
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    return np.concatenate(parts, axis=1)


def mode_match_distributed(eng, mode_id, dist, world):
    parts = [None] * world
    dist.all_gather_object(parts, eng.ReadModeMatchRaw(mode_id))
    return combine_mode_match(parts)


def steadystate_distributed(eng, period, dist, world):
    parts = [None] * world
    dist.all_gather_object(parts, eng.SteadyStateRaw())
    return combine_steadystate(parts, period)
