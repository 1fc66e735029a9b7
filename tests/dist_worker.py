"""Multi-rank worker (one process per rank) used by the distributed tests.

  backend "nccl": run under torchrun on >= 2 GPUs; every rank owns a slab of a global frame in
      librnmf_b200 (rnmf_frame_create_slab), runs halo-exchanging Jacobi sweeps and the result is
      compared, slab by slab, against the single-domain CPU oracle -- bit-exact.
  backend "gloo": CPU only (no GPU, no librnmf compute): the same decomposition rule
      (rnmf_slab_partition == Domain::new_rectangular's split) driven with the ORACLE as the
      per-slab sweep and torch.distributed(gloo) send/recv as the halo exchange; checks that slab
      frames + HALO faces + plane exchange reproduce the single-domain result exactly.
Exit code 0 = pass.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

MIXED_3D = ([("dirichlet", 0.0), ("neumann", 0.0), ("dirichlet", 0.25)],
            [("dirichlet", 1.0), ("neumann", -0.5), ("neumann", 0.5)])
MIXED_2D = ([("neumann", 0.3), ("dirichlet", -1.0)], [("dirichlet", 2.0), ("neumann", -0.7)])

CASES = [
    # n_global, hi, spec, sweeps, n_ghost
    ([40, 36, 50], [40.0, 72.0, 150.0], MIXED_3D, 5, 1),
    ([130, 20, 23], [1.0, 1.0, 1.0], MIXED_3D, 4, 1),      # uneven split: remainder goes to the last rank
    ([64, 37], [64.0, 37.0], MIXED_2D, 6, 1),
    ([20, 18, 29], [2.0, 2.0, 3.0], MIXED_3D, 3, 2),       # two ghost layers: stand-alone fill + 2-slice halos
    ([33, 41], [1.0, 2.0], MIXED_2D, 4, 2),
]
# extra cases for the temporally blocked kernels across slabs (3D: three sweeps per launch for even n0 / n1, otherwise two; deep halo:
# ghost + deep planes per neighbour)
TB2_CASES = [
    ([40, 36, 50], [40.0, 72.0, 150.0], MIXED_3D, 9, 1),   # 1 double | 3 doubles + the norm-carrying single
    ([130, 20, 23], [1.0, 1.0, 1.0], MIXED_3D, 8, 1),      # uneven split; 1 double | 2 doubles + 2 singles (world 8: slabs too thin, single sweeps)
    ([257, 30, 16], [3.0, 2.0, 1.0], MIXED_3D, 6, 1),      # 4-plane slabs at world 4 (the minimum), ragged x tiles
    ([65, 27, 43], [1.0, 2.0, 3.0], MIXED_3D, 7, 1),       # odd n0 / n1: x-hi / y-hi ghosts of the travelling planes; 5-plane slabs at world 8
    ([200, 40, 300], [2.0, 1.0, 3.0], MIXED_3D, 6, 1),     # slabs of several chunks (world 2: 150 planes = 96 + 54)
    ([260, 38, 96], [2.0, 1.0, 3.0], MIXED_3D, 12, 1),     # even n0 / n1: THREE sweeps per launch across slabs (12-plane slabs at world 8): a pair | 3 triples + the norm-carrying single
    ([128, 24, 48], [1.0, 1.0, 1.0], MIXED_3D, 7, 1),      # 6-plane slabs at world 8 (the minimum for three sweeps per launch): a pair | a triple, a single, the norm-carrying single
    # 2D y-slabs, four sweeps per launch (deep halo: ghost + 3 deep rows per neighbour)
    ([700, 130], [7.0, 1.3], MIXED_2D, 13, 1),             # two strips; 2 | 2 groups of four + 2 singles + the norm-carrying single
    ([1031, 67], [1.0, 1.0], MIXED_2D, 10, 1),             # odd n0; 8-row slabs at world 8 (the minimum)
    ([64, 400], [1.0, 4.0], MIXED_2D, 9, 1),               # several chunks per slab
]


def reference(n, hi, spec, sweeps, ng=1):
    from oracle import oracle as ora
    g = ora.geom(n, hi=hi, ng=ng)
    bc = ora.bcs(g.dim, [spec])
    nv = int(np.prod(n))
    psi, rhs = ora.zeros(g), ora.zeros(g)
    ora.valid_view(g, psi)[0] = ora.splitmix_field(nv, 101).reshape(tuple(reversed(n)))
    ora.valid_view(g, rhs)[0] = ora.splitmix_field(nv, 202).reshape(tuple(reversed(n)))
    ora.fill_bc(g, psi, bc)
    l1 = ora.jacobi(g, psi, rhs, bc, sweeps)
    return ora.valid_view(g, psi)[0].copy(), l1


def host_call_check(R, ctx, rank, world):
    """rnmf_jacobi_sweeps_host on slab frames (the per-rank end-to-end call): host slabs in, sweeps, host slab out.  With the
    default e2e_skew the upload is chunked and overlapped with band-wise double sweeps, the wedges next to the neighbours catch
    up afterwards (capi.cu: skewed_upload_and_first_sweeps); e2e_skew = 0 is the plain path.  Both must equal the oracle on the
    GLOBAL frame bit for bit.  The slab is just over the 64 MB threshold of the overlapped path."""
    import ctypes as C
    from oracle import oracle as ora
    from rnmf_b200 import _capi, poisson
    n, sweeps = [254, 126, 264 * world], 23
    hi = [float(v) for v in n]
    g = ora.geom(n, hi=hi)
    obc = ora.bcs(3, [MIXED_3D])
    rng = np.random.default_rng(5)
    psi0, rhs0 = rng.standard_normal(g.n_nodes), rng.standard_normal(g.n_nodes)
    want = psi0.copy()
    l1_ref = ora.jacobi(g, want, rhs0, obc, sweeps, omp=True)
    mk = {"dirichlet": R.BcType.Dirichlet, "neumann": R.BcType.Neumann}
    bc = R.BCs([R.ComponentBCs([mk[k](v) for k, v in MIXED_3D[0]], [mk[k](v) for k, v in MIXED_3D[1]])])
    mesh = R.CartesianMesh.new([0.0] * 3, hi, n)
    start, count = R.slab_partition(n[2], world, rank)
    plane = (n[0] + 2) * (n[1] + 2)
    # this rank's host frames: its slab of the global dense frames, ghost planes included (grown planes start .. start+count+1)
    sl = slice(plane * start, plane * (start + count + 2))
    psi_h, rhs_h = np.ascontiguousarray(psi0[sl]), np.ascontiguousarray(rhs0[sl])
    psi = R.CartesianDataFrame.new_from(mesh, bc, 1, 1, ctx, slab=True)
    rhs = R.CartesianDataFrame.new_from(mesh, bc, 1, 1, ctx, slab=True)
    psi.enable_peer_halo()
    coef = (C.c_double * 4)(*poisson.poisson_coefs(3, mesh.dx))
    ok = True
    for skew in (2, 0):
        ctx.set_option("e2e_skew", skew)
        ctx.set_option("e2e_skew_chunks", 4)
        ctx.set_option("e2e_skew_pairs", 9)
        out = np.empty_like(psi_h)
        l1 = C.c_double()
        k0 = ctx.kernel_launches
        _capi.check(_capi.lib().rnmf_jacobi_sweeps_host(psi._h, rhs._h, psi_h.ctypes.data, rhs_h.ctypes.data, coef, sweeps,
                                                        out.ctypes.data, C.byref(l1)))
        launches = ctx.kernel_launches - k0
        p0, p1 = (0 if rank == 0 else 1), (count + 2 if rank == world - 1 else count + 1)      # my valid planes + the outer ghost planes
        same = np.array_equal(out[plane * p0:plane * p1], want[plane * (start + p0):plane * (start + p1)])
        norm_ok = abs(l1.value - l1_ref) <= 1e-12 * abs(l1_ref)
        used = launches > 40 if skew else launches < 40          # the overlapped path issues one launch per (chunk, pair) band
        if not (same and norm_ok and used):
            ok = False
            print(f"[rank {rank}] FAIL host call e2e_skew={skew}: field={same} norm={norm_ok} ({l1.value} vs {l1_ref}) launches={launches}", flush=True)
    ctx.set_option("e2e_skew", 2)
    ctx.set_option("e2e_skew_chunks", 8)
    ctx.set_option("e2e_skew_pairs", 56)
    psi.close()
    rhs.close()
    return ok


def run_nccl():
    import torch
    import torch.distributed as dist
    import rnmf_b200 as R
    from rnmf_b200 import poisson
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = R.Context.from_torch_distributed()
    mk = {"dirichlet": R.BcType.Dirichlet, "neumann": R.BcType.Neumann}
    ok = True
    modes = os.environ.get("RNMF_DIST_MODES", "peer,peer_tb2,peer_flag_kernels,nccl_overlap,nccl_serial").split(",")
    for mode in modes:
        ctx.set_option("halo_overlap", 0 if mode == "nccl_serial" else 1)
        ctx.set_option("peer_halo", 1 if mode.startswith("peer") else 0)
        ctx.set_option("peer_inkernel", 1 if mode in ("peer", "peer_tb2") else 0)   # handshake inside the sweep kernel vs 1-thread flag kernels
        overlap = mode
        for n, hi, spec, sweeps, ng in CASES + (TB2_CASES if mode == "peer_tb2" else []):
            dim = len(n)
            # "peer_tb2" = the defaults (3D: two sweeps per launch across slabs); the other modes: one sweep per launch only
            ctx.set_option("sweep_variant", "auto" if mode == "peer_tb2" else (6 if dim == 3 else 20))
            want, l1_ref = reference(n, hi, spec, sweeps, ng)
            bc = R.BCs([R.ComponentBCs([mk[k](v) for k, v in spec[0]], [mk[k](v) for k, v in spec[1]])])
            mesh = R.CartesianMesh.new([0.0] * dim, hi, n)
            start, count = R.slab_partition(n[-1], world, rank)
            psi = R.CartesianDataFrame.new_from(mesh, bc, 1, ng, ctx, slab=True)
            rhs = R.CartesianDataFrame.new_from(mesh, bc, 1, ng, ctx, slab=True)
            psi.fill_synthetic(101)
            rhs.fill_synthetic(202)
            psi.fill_bc()
            if mode.startswith("peer") and ng == 1:
                psi.enable_peer_halo()
                # two calls: the epoch handshake and buffer parity must carry over between calls
                before = ctx.kernel_launches
                poisson.jacobi_sweeps(psi, rhs, 2)
                if mode == "peer_tb2" and dim == 3 and n[-1] // world >= 4:
                    # shell copy + ONE double-sweep launch + the closing wait kernel
                    if ctx.kernel_launches - before != 3:
                        ok = False
                        print(f"[rank {rank}] FAIL n={n}: double-sweep kernel not used ({ctx.kernel_launches - before} launches)", flush=True)
                l1 = poisson.jacobi_sweeps(psi, rhs, sweeps - 2, want_norm=True)
            else:
                l1 = poisson.jacobi_sweeps(psi, rhs, sweeps, want_norm=True)
            got = psi.get_valid_cells().reshape((count,) + tuple(reversed(n[:-1])))
            same = np.array_equal(got, want[start:start + count])
            norm_ok = abs(l1 - l1_ref) <= 1e-12 * abs(l1_ref)
            total = poisson.sum(psi)
            sum_ok = abs(total - np.abs(want).sum()) <= 1e-12 * np.abs(want).sum()
            if not (same and norm_ok and sum_ok):
                ok = False
                print(f"[rank {rank}] FAIL n={n} overlap={overlap}: field={same} norm={norm_ok} ({l1} vs {l1_ref}) sum={sum_ok}",
                      flush=True)
            psi.close()
            rhs.close()
    ctx.set_option("sweep_variant", "auto")
    ctx.set_option("peer_halo", 1)
    ctx.set_option("peer_inkernel", 1)
    ctx.set_option("halo_overlap", 1)
    ok = host_call_check(R, ctx, rank, world) and ok
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    ctx.close()
    dist.destroy_process_group()
    if rank == 0:
        print("dist nccl parity:", "PASS" if flag.item() == 0 else "FAIL", f"(world={world})", flush=True)
    return int(flag.item() != 0)


def run_gloo(rank: int, world: int, port: int, results):
    """CPU: oracle per slab + gloo halo exchange == oracle on the whole domain."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from oracle import oracle as ora
    from rnmf_b200 import slab_partition
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    for n, hi, spec, sweeps, ng in CASES:
        dim = len(n)
        D = dim - 1
        want, _ = reference(n, hi, spec, sweeps, ng)
        start, count = slab_partition(n[D], world, rank)
        dx = [hi[d] / n[d] for d in range(dim)]
        nl = list(n)
        nl[D] = count
        lo = [0.0] * dim
        lo[D] = start * dx[D]
        hil = list(hi)
        hil[D] = lo[D] + count * dx[D]
        g = ora.geom(nl, lo=lo, hi=hil, ng=ng)
        for d in range(dim):
            g.dx[d] = dx[d]                       # exactly the global dx (what the C ABI passes)
        lo_spec, hi_spec = list(spec[0]), list(spec[1])
        if rank > 0:
            lo_spec[D] = ("halo", None)
        if rank < world - 1:
            hi_spec[D] = ("halo", None)
        bc = ora.bcs(dim, [(lo_spec, hi_spec)])
        nv_glob = int(np.prod(n))
        per_slice = int(np.prod(n[:D]))
        psi, rhs = ora.zeros(g), ora.zeros(g)
        full_psi = ora.splitmix_field(nv_glob, 101).reshape(tuple(reversed(n)))
        full_rhs = ora.splitmix_field(nv_glob, 202).reshape(tuple(reversed(n)))
        ora.valid_view(g, psi)[0] = full_psi[start:start + count]
        ora.valid_view(g, rhs)[0] = full_rhs[start:start + count]
        ora.fill_bc(g, psi, bc)
        grid = ora.as_grid(g, psi)[0]

        def exchange():
            # first/last valid slices -> neighbours' ghost slices (whole grown slices, as the library does)
            reqs = []
            recv_lo = torch.empty(grid[:ng].shape, dtype=torch.float64)
            recv_hi = torch.empty(grid[:ng].shape, dtype=torch.float64)
            if rank > 0:
                reqs.append(dist.isend(torch.from_numpy(grid[ng:2 * ng].copy()), rank - 1))
                reqs.append(dist.irecv(recv_lo, rank - 1))
            if rank < world - 1:
                reqs.append(dist.isend(torch.from_numpy(grid[-2 * ng:-ng].copy()), rank + 1))
                reqs.append(dist.irecv(recv_hi, rank + 1))
            for r in reqs:
                r.wait()
            if rank > 0:
                grid[:ng] = recv_lo.numpy()
            if rank < world - 1:
                grid[-ng:] = recv_hi.numpy()

        exchange()
        for _ in range(sweeps):
            ora.jacobi(g, psi, rhs, bc, 1)
            exchange()
        got = ora.valid_view(g, psi)[0]
        if not np.array_equal(got, want[start:start + count]):
            ok = False
            print(f"[gloo rank {rank}] FAIL n={n}", flush=True)
        del per_slice
    t = torch.tensor([0 if ok else 1])
    dist.all_reduce(t)
    dist.destroy_process_group()
    if results is not None:
        results[rank] = int(t.item())
    return int(t.item() != 0)


if __name__ == "__main__":
    backend = sys.argv[1] if len(sys.argv) > 1 else "nccl"
    if backend == "nccl":
        sys.exit(run_nccl())
    else:
        sys.exit(run_gloo(int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("MASTER_PORT", "29533")), None))
