"""CPU: domain decomposition (host logic) pinned bit-exactly by the loop oracle, and the N>1 path
exercised with world_size-2 gloo processes (the per-rank compute is the numpy oracle here -- the
product's compute needs a GPU and is covered by the -m gpu tests)."""
import os
import socket

import numpy as np
import pytest

import moka_oracle as O
import partition_oracle as PO
from conftest import hex_mesh
from moka_b200 import partition


@pytest.mark.parametrize("nx,ny,nparts", [(16, 16, 2), (16, 12, 3), (24, 16, 4), (16, 16, 8)])
def test_partition_and_halo_lists_bit_exact_vs_loop_oracle(nx, ny, nparts):
    m = hex_mesh(nx, ny, 1000.0)
    part = partition.rcb_partition(m["xCell"], m["yCell"], nparts)
    part_o = PO.rcb_partition(m["xCell"].tolist(), m["yCell"].tolist(), nparts)
    assert part.dtype == np.int32 and part.tolist() == part_o
    sizes = np.bincount(part, minlength=nparts)
    assert sizes.max() - sizes.min() <= nparts                    # balanced
    locs = partition.decompose(m, nparts, part)
    mo = {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in m.items()
          if k in ("nCells", "cellsOnEdge", "edgesOnCell", "nEdgesOnCell")}
    sets, halos = PO.halo_lists(mo, part_o, nparts)
    owned_cells = np.zeros(m["nCells"], int)
    owned_edges = np.zeros(m["nEdges"], int)
    for r, loc in enumerate(locs):
        cells, nco, edges, neo = sets[r]
        assert loc["cellsGlobal"].tolist() == cells and loc["nCellsOwned"] == nco
        assert loc["edgesGlobal"].tolist() == edges and loc["nEdgesOwned"] == neo
        owned_cells[loc["cellsGlobal"][:nco]] += 1
        owned_edges[loc["edgesGlobal"][:neo]] += 1
        assert sorted(loc["halo"]["recv"]) == sorted(q for q in halos[r]["recv"])
        for q in loc["halo"]["peers"]:
            assert loc["halo"]["recv"][q].tolist() == halos[r]["recv"].get(q, [])
            assert loc["halo"]["send"][q].tolist() == halos[r]["send"].get(q, [])
        # every stencil of an owned entity stays inside the local mesh
        no, ne = loc["nCellsOwned"], loc["nEdgesOwned"]
        assert np.all(loc["cellsOnEdge"][:ne] > 0)
        assert np.all(loc["edgesOnEdge"][:ne] > 0)
        assert np.all(loc["edgesOnCell"][:no] > 0)
        nb = loc["cellsOnEdge"][loc["edgesOnCell"][:no].astype(np.int64) - 1]
        assert np.all(nb > 0)
    assert np.all(owned_cells == 1) and np.all(owned_edges == 1)   # a partition: every entity owned exactly once


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nx, nsteps, out_dir):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [here, os.path.join(os.path.dirname(here), "mpas-ocean.jl_b200"), os.path.join(os.path.dirname(here), "oracle")]
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import moka_b200.planar_hex as ph
    import moka_oracle_c as OC
    from moka_b200 import multi_gpu
    m = ph.periodic_hex(nx, nx, 1.0e7 / nx, with_dual=False)
    OC.sign_index_fields(m)
    ssh, u, h = O.InertialGravityWave(m).initial_state()
    loc = partition.decompose(m, world)[rank]
    OC.sign_index_fields(loc)
    sidx, scnt, ridx, rcnt = partition.flat_halo(loc, world)
    ex = multi_gpu.HaloExchanger(scnt, rcnt, torch.float64, "cpu")
    nCl = loc["nCells"]
    ls, lu, lh = multi_gpu.local_state(loc, ssh, u, h)
    dt = 0.5 * (1.0e7 / nx) / np.sqrt(O.GRAVITY * 1000.0)
    a, b = [dt / 2, dt / 2, dt], [dt / 6, dt / 3, dt / 3, dt / 6]

    def exchange(uu, hh):
        comb = np.concatenate([hh, uu])
        ex.send[:len(sidx)] = torch.from_numpy(comb[sidx])
        ex.exchange()
        comb[ridx] = ex.recv[:len(ridx)].numpy()
        return comb[nCl:], comb[:nCl]

    u_cur, h_cur = lu.copy(), lh.copy()
    for _ in range(nsteps):
        u_pro, h_pro, u_new, h_new = u_cur.copy(), h_cur.copy(), u_cur.copy(), h_cur.copy()
        for s in range(4):
            tu, th = O.tendencies_consistent(loc, u_pro, h_pro)
            if s < 3:
                u_pro, h_pro = exchange(u_cur + a[s] * tu, h_cur + a[s] * th)
            u_new, h_new = u_new + b[s] * tu, h_new + b[s] * th
        u_cur, h_cur = exchange(u_new, h_new)
    no, ne = loc["nCellsOwned"], loc["nEdgesOwned"]
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), u=u_cur[:ne], h=h_cur[:no], ce=loc["cellsGlobal"][:no], ee=loc["edgesGlobal"][:ne],
             halo_u=u_cur[ne:], halo_e=loc["edgesGlobal"][ne:])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_two_rank_gloo_halo_exchange_matches_single_domain(tmp_path, world):
    import torch.multiprocessing as mp
    nx, nsteps = 16, 5
    mp.spawn(_worker, args=(world, _free_port(), nx, nsteps, str(tmp_path)), nprocs=world, join=True)
    m = hex_mesh(nx, with_dual=False) if False else hex_mesh(nx)
    ssh, u, h = O.InertialGravityWave(m).initial_state()
    prog = O.new_state(m, ssh, u, h)
    dt = 0.5 * (1.0e7 / nx) / np.sqrt(O.GRAVITY * 1000.0)
    for _ in range(nsteps):
        O.timestep_rk4(m, prog, dt)
    gu, gh = np.full(m["nEdges"], np.nan), np.full(m["nCells"], np.nan)
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        gu[z["ee"]], gh[z["ce"]] = z["u"], z["h"]
        assert np.array_equal(z["halo_u"], prog["normalVelocity"][-1][z["halo_e"]])     # halos hold the owners' values
    assert np.array_equal(gu, prog["normalVelocity"][-1])         # bit-exact: same arithmetic per entity
    assert np.array_equal(gh, prog["layerThickness"][-1])


def _worker_fe(rank, world, port, nx, nsteps, out_dir):
    """ForwardEuler over two gloo ranks with the numpy oracle as the per-rank compute: the halo copies of (h, u) and of
    (ssh, layerThicknessEdge) after every step are all a rank needs -- first-step quirk and lagged
    layerThicknessEdge included."""
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [here, os.path.join(os.path.dirname(here), "mpas-ocean.jl_b200"), os.path.join(os.path.dirname(here), "oracle")]
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import moka_b200.planar_hex as ph
    import moka_oracle_c as OC
    from moka_b200 import multi_gpu
    m = ph.periodic_hex(nx, nx, 1.0e7 / nx, with_dual=False)
    OC.sign_index_fields(m)
    ssh, u, h = O.InertialGravityWave(m).initial_state()
    loc = partition.decompose(m, world)[rank]
    OC.sign_index_fields(loc)
    sidx, scnt, ridx, rcnt = partition.flat_halo(loc, world)
    ex = multi_gpu.HaloExchanger(scnt, rcnt, torch.float64, "cpu")
    nCl = loc["nCells"]
    ls, lu, lh = multi_gpu.local_state(loc, ssh, u, h)
    dt = 0.5 * (1.0e7 / nx) / np.sqrt(O.GRAVITY * 1000.0)

    def exchange(edge_values, cell_values):
        comb = np.concatenate([cell_values, edge_values])
        ex.send[:len(sidx)] = torch.from_numpy(comb[sidx])
        ex.exchange()
        comb[ridx] = ex.recv[:len(ridx)].numpy()
        return comb[nCl:], comb[:nCl]

    hE = np.zeros(loc["nEdges"])                                          # DiagnosticVars start at zero: no flux in the first step
    H = O.resting_thickness_sum(loc)
    for _ in range(nsteps):
        flux = lu * hE                                                   # the lagged layerThicknessEdge (DiagnosticVars.jl:141-173)
        hE_new = O.interpolate_cell2edge(loc, lh)
        tu = O.compute_normal_velocity_tendency(loc, ls, lu)
        th = O.compute_layer_thickness_tendency(loc, flux)
        lu, lh = exchange(lu + dt * tu, lh + dt * th)
        hE, ls = exchange(hE_new, lh - H)
    no, ne = loc["nCellsOwned"], loc["nEdgesOwned"]
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), u=lu[:ne], h=lh[:no], ssh=ls[:no], ce=loc["cellsGlobal"][:no], ee=loc["edgesGlobal"][:ne])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_forward_euler_matches_single_domain(tmp_path):
    import torch.multiprocessing as mp
    nx, nsteps, world = 16, 6, 2
    mp.spawn(_worker_fe, args=(world, _free_port(), nx, nsteps, str(tmp_path)), nprocs=world, join=True)
    m = hex_mesh(nx)
    ssh, u, h = O.InertialGravityWave(m).initial_state()
    prog, diag = O.new_state(m, ssh, u, h), O.new_diag(m)
    dt = 0.5 * (1.0e7 / nx) / np.sqrt(O.GRAVITY * 1000.0)
    for _ in range(nsteps):
        O.timestep_forward_euler(m, prog, diag, dt)
    gu, gh, gs = np.full(m["nEdges"], np.nan), np.full(m["nCells"], np.nan), np.full(m["nCells"], np.nan)
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        gu[z["ee"]], gh[z["ce"]], gs[z["ce"]] = z["u"], z["h"], z["ssh"]
    assert np.array_equal(gu, prog["normalVelocity"][-1])         # bit-exact: same arithmetic per entity
    assert np.array_equal(gh, prog["layerThickness"][-1])
    assert np.array_equal(gs, prog["ssh"][-1])


def _levels_case(nx, K):
    m = dict(hex_mesh(nx, with_dual=False))
    ssh, u, h = O.InertialGravityWave(m).initial_state()
    frac = np.random.default_rng(K).uniform(0.5, 1.5, K)
    frac /= frac.sum()
    rest = np.outer(np.full(m["nCells"], 1000.0), frac)
    m["restingThickness"], m["nVertLevels"] = rest, K
    return m, np.ascontiguousarray(np.outer(u, 1.0 + 0.1 * np.arange(K)).T), np.ascontiguousarray((rest + np.outer(ssh, frac)).T)


def _worker_levels(rank, world, port, nx, K, nsteps, out_dir):
    """RungeKutta4 of a K-level state over gloo ranks with the numpy oracle (level axis by broadcasting) as the per-rank compute:
    the halo copies of every level of (h, u) after each stage are all a rank needs -- the free surface of a halo cell follows
    from its column (the library sends it as one more plane because its kernels form ssh for owned cells only)."""
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [here, os.path.join(os.path.dirname(here), "mpas-ocean.jl_b200"), os.path.join(os.path.dirname(here), "oracle")]
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import moka_oracle_c as OC
    from moka_b200 import multi_gpu
    m, uk, hk = _levels_case(nx, K)
    OC.sign_index_fields(m)
    loc = partition.decompose(m, world)[rank]
    OC.sign_index_fields(loc)
    sidx, scnt, ridx, rcnt = partition.flat_halo(loc, world)
    ex = multi_gpu.HaloExchanger(scnt, rcnt, torch.float64, "cpu")
    nCl, no, ne = loc["nCells"], loc["nCellsOwned"], loc["nEdgesOwned"]
    dt = 0.5 * (1.0e7 / nx) / np.sqrt(O.GRAVITY * 1000.0)
    a, b = [dt / 2, dt / 2, dt], [dt / 6, dt / 3, dt / 3, dt / 6]

    def exchange(uu, hh):
        uo, ho = np.empty_like(uu), np.empty_like(hh)
        for k in range(K):
            comb = np.concatenate([hh[k], uu[k]])
            ex.send[:len(sidx)] = torch.from_numpy(comb[sidx])
            ex.exchange()
            comb[ridx] = ex.recv[:len(ridx)].numpy()
            uo[k], ho[k] = comb[nCl:], comb[:nCl]
        return uo, ho

    u_cur, h_cur = uk[:, loc["edgesGlobal"]].copy(), hk[:, loc["cellsGlobal"]].copy()
    for _ in range(nsteps):
        u_pro, h_pro, u_new, h_new = u_cur.copy(), h_cur.copy(), u_cur.copy(), h_cur.copy()
        for s in range(4):
            tu, th = O.tendencies_consistent(loc, u_pro, h_pro)
            if s < 3:
                u_pro, h_pro = exchange(u_cur + a[s] * tu, h_cur + a[s] * th)
            u_new, h_new = u_new + b[s] * tu, h_new + b[s] * th
        u_cur, h_cur = exchange(u_new, h_new)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), u=u_cur[:, :ne], h=h_cur[:, :no], ce=loc["cellsGlobal"][:no], ee=loc["edgesGlobal"][:ne])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_multilevel_matches_single_domain(tmp_path):
    import torch.multiprocessing as mp
    nx, K, nsteps, world = 16, 3, 3, 2
    mp.spawn(_worker_levels, args=(world, _free_port(), nx, K, nsteps, str(tmp_path)), nprocs=world, join=True)
    m, uk, hk = _levels_case(nx, K)
    prog = O.new_state(m, np.zeros(m["nCells"]), uk, hk)
    dt = 0.5 * (1.0e7 / nx) / np.sqrt(O.GRAVITY * 1000.0)
    for _ in range(nsteps):
        O.timestep_rk4(m, prog, dt)
    gu, gh = np.full((K, m["nEdges"]), np.nan), np.full((K, m["nCells"]), np.nan)
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        gu[:, z["ee"]], gh[:, z["ce"]] = z["u"], z["h"]
    assert np.array_equal(gu, prog["normalVelocity"][-1]) and np.array_equal(gh, prog["layerThickness"][-1])     # same arithmetic per entity


def _worker_reverse(rank, world, port, nx, nsteps, out_dir):
    """The reverse sweep of RungeKutta4 over gloo ranks with the scatter-form adjoint oracle as the per-rank compute: plain halo
    copies -- over the SAME send / receive lists as the forward exchange -- of the stage states, of kbar after every reversed
    stage and of lambda after every reversed step are all a rank needs for the gradient on its owned entities (DESIGN.md
    section 7; what mokab_adjoint_rk4 does on a decomposed state, independent of the CUDA code)."""
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [here, os.path.join(os.path.dirname(here), "mpas-ocean.jl_b200"), os.path.join(os.path.dirname(here), "oracle")]
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import adjoint_oracle as A
    import moka_b200.planar_hex as ph
    import moka_oracle_c as OC
    from moka_b200 import multi_gpu
    m = ph.periodic_hex(nx, nx, 1.0e7 / nx, with_dual=False)
    OC.sign_index_fields(m)
    ssh, u, h = O.InertialGravityWave(m).initial_state()
    loc = partition.decompose(m, world)[rank]
    OC.sign_index_fields(loc)
    sidx, scnt, ridx, rcnt = partition.flat_halo(loc, world)
    ex = multi_gpu.HaloExchanger(scnt, rcnt, torch.float64, "cpu")
    nCl, no, ne = loc["nCells"], loc["nCellsOwned"], loc["nEdgesOwned"]
    dt = 0.5 * (1.0e7 / nx) / np.sqrt(O.GRAVITY * 1000.0)
    a, b = [dt / 2, dt / 2, dt], [dt / 6, dt / 3, dt / 3, dt / 6]

    def exchange(uu, hh):
        comb = np.concatenate([hh, uu])
        comb[nCl + ne:] = np.nan                                          # what a rank does not own is whatever the exchange brings
        comb[no:nCl] = np.nan
        ex.send[:len(sidx)] = torch.from_numpy(comb[sidx])
        ex.exchange()
        comb[ridx] = ex.recv[:len(ridx)].numpy()
        return comb[nCl:], comb[:nCl]

    # the forward trajectory (not under test here: tests above) from the single-domain oracle, restricted to the local mesh
    traj = A.run_forward(m, u, h, dt, nsteps)
    lam_u = np.zeros(loc["nEdges"])
    lam_h = 2.0 * (traj[-1][1] - O.resting_thickness_sum(m))[loc["cellsGlobal"]]
    for n in range(nsteps - 1, -1, -1):
        ys = [(yu[loc["edgesGlobal"]], yh[loc["cellsGlobal"]]) for yu, yh in A.rk4_stage_states(m, traj[n][0], traj[n][1], dt)]
        out_u, out_h = lam_u.copy(), lam_h.copy()
        kbu, kbh = b[3] * lam_u, b[3] * lam_h
        for s in (3, 2, 1, 0):
            ybu, ybh = A.tendencies_vjp(loc, ys[s][0], ys[s][1], kbu, kbh)
            out_u, out_h = out_u + ybu, out_h + ybh
            if s > 0:
                kbu, kbh = exchange(b[s - 1] * lam_u + a[s - 1] * ybu, b[s - 1] * lam_h + a[s - 1] * ybh)
        lam_u, lam_h = exchange(out_u, out_h)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), gu=lam_u[:ne], gh=lam_h[:no], ce=loc["cellsGlobal"][:no], ee=loc["edgesGlobal"][:ne])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_reverse_sweep_with_plain_halo_copies_matches_single_domain(tmp_path, world):
    import adjoint_oracle as A
    import torch.multiprocessing as mp
    nx, nsteps = 16, 3
    mp.spawn(_worker_reverse, args=(world, _free_port(), nx, nsteps, str(tmp_path)), nprocs=world, join=True)
    m = hex_mesh(nx)
    ssh, u, h = O.InertialGravityWave(m).initial_state()
    dt = 0.5 * (1.0e7 / nx) / np.sqrt(O.GRAVITY * 1000.0)
    _, gu, gh = A.gradient_sum_ssh2(m, u, h, dt, nsteps)
    au, ah = np.full(m["nEdges"], np.nan), np.full(m["nCells"], np.nan)
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        au[z["ee"]], ah[z["ce"]] = z["gu"], z["gh"]
    assert np.linalg.norm(au - gu) <= 1e-12 * np.linalg.norm(gu)          # (sums reassociated: not bit for bit)
    assert np.linalg.norm(ah - gh) <= 1e-12 * np.linalg.norm(gh)


def test_as_many_parts_as_cells_and_one_more():
    """One cell per part is the limit: every rank must own a cell (an empty rank would launch empty grids); beyond it the
    partitioner refuses with a message instead of failing somewhere inside numpy."""
    m = hex_mesh(4)
    locs = partition.decompose(m, m["nCells"])
    assert sorted(int(loc["cellsGlobal"][0]) for loc in locs) == list(range(m["nCells"]))
    assert all(loc["nCellsOwned"] == 1 for loc in locs)
    with pytest.raises(ValueError, match="must be between 1 and the number of cells"):
        partition.decompose(m, m["nCells"] + 1)
    with pytest.raises(ValueError, match="must be between 1"):
        partition.rcb_partition(m["xCell"], m["yCell"], 0)


def _exchange_worker(rank, world, port, out_dir):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [here, os.path.join(os.path.dirname(here), "mpas-ocean.jl_b200")]
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from moka_b200 import multi_gpu
    # the control plane of a decomposed run: rank 0's 128-byte communicator id reaches every rank, through a process group ...
    rt = multi_gpu.TorchRuntime(0, device="cpu")
    blob = bytes(range(128)) if rank == 0 else None
    ok = rt.broadcast_bytes(blob, 0) == bytes(range(128)) and rt.rank_and_size() == (rank, world)
    # ... or through a bare TCP store (no process group: what a Julia host would do with MPI.jl or a file)
    st = multi_gpu.StoreRuntime(rank, world, "127.0.0.1", port + 1)
    ok = ok and st.broadcast_bytes(bytes([7]) * 128 if rank == 0 else None, 0) == bytes([7]) * 128
    ok = ok and st.broadcast_bytes(bytes([9]) * 128 if rank == 0 else None, 0) == bytes([9]) * 128 and st.rank_and_size() == (rank, world)
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_runtime_host_exchanges_over_gloo(tmp_path):
    """What is left of the host side of a decomposed run -- handing rank 0's communicator id to the others -- on 3 ranks
    (everything else, the set-up exchanges of the direct-store path included, is inside the library: csrc/decomposed.cuh,
    exercised with emulated ranks by tests/sim/check_decomposed.py)."""
    import torch.multiprocessing as mp
    world = 3
    mp.spawn(_exchange_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert [open(tmp_path / f"ok{r}").read() for r in range(world)] == ["1"] * world


def test_decompose_one_equals_the_all_ranks_decomposition():
    """partition.decompose_one (what every rank of the bench computes for itself) against decompose (all ranks at once), on a
    hexagon mesh, a channel and a Voronoi mesh, for part counts that produce corner-only neighbours."""
    import moka_b200 as mb
    from moka_b200.planar_voronoi import periodic_voronoi
    meshes = [mb.periodic_hex(24, 24, 1.0e5, with_dual=False), mb.channel_hex(20, 20, 1.0e5),
              periodic_voronoi(16, 16, 1.0e5, jitter=0.3, seed=4)]
    for m in meshes:
        for P in (2, 3, 4, 8):
            locs = partition.decompose(m, P)
            sets = [(loc["cellsGlobal"][loc["nCellsOwned"]:], loc["edgesGlobal"][loc["nEdgesOwned"]:]) for loc in locs]
            for r in range(P):
                one = partition.decompose_one(m, P, r) if r % 2 else partition.decompose_one(m, P, r, halo_sets=sets)
                ref = locs[r]
                assert one["halo"]["peers"] == ref["halo"]["peers"]
                for q in ref["halo"]["peers"]:
                    assert np.array_equal(one["halo"]["send"][q], ref["halo"]["send"][q]) and np.array_equal(one["halo"]["recv"][q], ref["halo"]["recv"][q])
                for k, v in ref.items():
                    if isinstance(v, np.ndarray):
                        assert np.array_equal(one[k], v), k
                assert (one["rank"], one["nparts"], one["nCellsOwned"], one["nEdgesOwned"]) == (r, P, ref["nCellsOwned"], ref["nEdgesOwned"])
