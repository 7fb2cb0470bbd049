"""Randomised sweep of the library on the simulated runtime: random meshes (regular hexagons of odd shapes, channel meshes with
wall masks, periodic Voronoi meshes from squares to octagons, spheres), random decompositions (2..9 ranks, all three halo
paths; RungeKutta4 and, on the packed path, the staged ForwardEuler), every stepper and both adjoints, a random scheduling policy -- each compared with the CPU oracle (bit for bit where
the operation order is the reference's, rel-L2 <= 1e-12 where weights are folded or sums reassociated).  Test infrastructure.

  python tests/sim/fuzz.py [--iterations 30] [--seed 0]
Prints one line per case and FUZZ_OK / FUZZ_FAILED."""
import argparse
import ctypes
import os
import sys
import time
import traceback

os.environ.setdefault("OMP_NUM_THREADS", "2")

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "mpas-ocean.jl_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), HERE]
os.environ["MOKAB_SIM"] = "1"
import simcuda  # noqa: E402
from moka_b200 import _lib  # noqa: E402

simcuda.runtime()
_lib.bind(ctypes.CDLL(simcuda._build.LIB))
import adjoint_oracle as A  # noqa: E402
import moka_b200 as mb  # noqa: E402
import moka_oracle_c as OC  # noqa: E402
from moka_b200 import multi_gpu, partition  # noqa: E402
from test_gpu_decomposed import _run_emulated, _run_emulated_fe  # noqa: E402

G, DEPTH = 9.80616, 1000.0


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-300))


def make_case(rng):
    kind = rng.choice(["hex", "channel", "voronoi", "voronoi", "sphere"])
    if kind == "hex":
        nx, ny = int(rng.integers(4, 28)), 2 * int(rng.integers(2, 14))
        m = mb.periodic_hex(nx, ny, 1.0e7 / nx)
        mo = m
        state = mb.inertialGravityWave(m).initial_state()
        dt, desc = mb.cfl_dt(m["dc"]), f"hex {nx}x{ny}"
    elif kind == "channel":
        nx = 2 * int(rng.integers(3, 12))
        m = mb.channel_hex(nx, nx, 1.0e7 / nx)
        mo = OC.apply_boundary_mask(m)
        state = mb.kelvinWave(m).initial_state()
        dt, desc = mb.cfl_dt(m["dc"]), f"channel {nx}x{nx}"
    elif kind == "voronoi":
        nx, ny = int(rng.integers(6, 26)), 2 * int(rng.integers(3, 13))
        jitter, seed = float(rng.uniform(0.0, 0.34)), int(rng.integers(0, 1000))
        m = mb.periodic_voronoi(nx, ny, 1.0e7 / nx, jitter=jitter, seed=seed, allow_obtuse=True)
        mo = m
        state = mb.inertialGravityWave(m).initial_state()
        dt, desc = 0.2 * mb.cfl_dt(m["dc"]), f"voronoi {nx}x{ny} jitter {jitter:.2f} seed {seed} polygons {np.bincount(m['nEdgesOnCell'])[3:].tolist()}"
    else:
        n = int(rng.integers(40, 900))
        m = mb.spherical_voronoi(n)
        mo = m
        ssh, u, h = mb.geostrophic_zonal_flow(m)
        state = (ssh, u + 0.3 * rng.standard_normal(m["nEdges"]), h)
        dt, desc = 0.25 * float(m["dcEdge"].min()) / float(np.sqrt(G * DEPTH)), f"sphere {n} polygons {np.bincount(m['nEdgesOnCell'])[3:].tolist()}"
    OC.sign_index_fields(mo)
    return kind, m, mo, state, dt, desc


def one_case(rng, backend):
    kind, m, mo, (ssh, u, h), dt, desc = make_case(rng)
    trace = (lambda *a: print("   ..", *a, flush=True)) if os.environ.get("FUZZ_TRACE") else (lambda *a: None)
    trace(desc)
    policy = str(rng.choice(["fifo", "lazy", "others_first", "random"]))
    simcuda.set_policy(policy, int(rng.integers(1, 1 << 30)))
    uniform_f = float(np.ptp(m["fEdge"])) == 0.0
    renumber = bool(rng.integers(0, 2))
    nsteps = int(rng.integers(1, 9))
    mesh = mb.Mesh(m, backend, renumber=renumber, explicit_eoe=bool(rng.integers(0, 2)), keep_widths=bool(rng.integers(0, 4) == 0))
    problems = []
    # RungeKutta4, fused and reference-order
    om = OC.OracleModel(mo, ssh, u, h)
    om.run_loop(dt, nsteps, "RungeKutta4")
    prog = mb.PrognosticVars(ssh, u, h, 2, mesh)
    mb.ocn_timestep(dt, prog, None, None, None, mb.RungeKutta4, nsteps=nsteps)
    if uniform_f:
        ok = np.array_equal(prog.normalVelocity, om.normalVelocity[1]) and np.array_equal(prog.layerThickness, om.layerThickness[1])
    else:
        ok = rel(prog.normalVelocity, om.normalVelocity[1]) <= 1e-12 and rel(prog.layerThickness, om.layerThickness[1]) <= 1e-12
    problems += [] if ok else ["fused RK4"]
    unf = mb.PrognosticVars(ssh, u, h, 2, mesh)
    mb.ocn_timestep(dt, unf, None, None, None, mb.RungeKutta4, nsteps=nsteps, fused=False)
    if not (np.array_equal(unf.normalVelocity, om.normalVelocity[1]) and np.array_equal(unf.layerThickness, om.layerThickness[1])):
        problems.append("unfused RK4")
    # ForwardEuler (fused where the widths allow)
    ofe = OC.OracleModel(mo, ssh, u, h)
    ofe.run_loop(dt, nsteps, "ForwardEuler")
    pfe = mb.PrognosticVars(ssh, u, h, 2, mesh)
    mb.ocn_timestep(dt, pfe, None, None, None, mb.ForwardEuler, nsteps=nsteps)
    if not (np.array_equal(pfe.normalVelocity, ofe.normalVelocity[1]) and np.array_equal(pfe.layerThickness, ofe.layerThickness[1])
            and np.array_equal(pfe.ssh, ofe.ssh[1])):
        problems.append("ForwardEuler")
    # both adjoints
    trace("adjoints")
    na = min(nsteps, 4)
    pa = mb.PrognosticVars(ssh, u, h, 2, mesh)
    d = mb.ocn_init_shadows(pa)
    mb.autodiff_reverse_run_loop(dt, pa, d, None, None, None, mb.RungeKutta4, na)
    _, gu, gh = A.gradient_sum_ssh2(mo, u, h, dt, na)
    if not (rel(d.normalVelocity, gu) <= 1e-11 and rel(d.layerThickness, gh) <= 1e-11):
        problems.append("RK4 adjoint")
    pa = mb.PrognosticVars(ssh, u, h, 2, mesh)
    d = mb.ocn_init_shadows(pa)
    mb.autodiff_reverse_run_loop(dt, pa, d, None, None, None, mb.ForwardEuler, na)
    _, gu, gh, gs, _ = A.gradient_sum_ssh2_fe(mo, ssh, u, h, dt, na)
    if not (rel(d.normalVelocity, gu) <= 1e-11 and rel(d.layerThickness, gh) <= 1e-11 and rel(d.ssh, gs) <= 1e-11):
        problems.append("ForwardEuler adjoint")
    # decomposition with ranks emulated in this process: the same bits as the single-domain fused run
    nparts = int(rng.integers(2, 10))
    halo = str(rng.choice(["nccl", "p2p", "p2p_fused"]))
    trace("decomposition", nparts, halo)
    if m["nCells"] >= 12 * nparts:
        md = {k: v for k, v in m.items() if k not in ("edgesOnVertex", "cellsOnVertex", "verticesOnEdge", "kiteAreasOnVertex",
                                                      "areaTriangle", "verticesOnCell", "edgeSignOnVertex")}
        md["nVertices"] = 0
        gu, gh, gs, _ = _run_emulated(backend, md, (ssh, u, h), nparts, dt, nsteps, halo=halo)
        base = mb.PrognosticVars(ssh, u, h, 2, mb.Mesh(m, backend))
        mb.ocn_timestep(dt, base, None, None, None, mb.RungeKutta4, nsteps=nsteps)
        if not (np.array_equal(gu, base.normalVelocity) and np.array_equal(gh, base.layerThickness)):
            problems.append(f"decomposed {nparts} ranks {halo}")
        desc += f"; {nparts} ranks {halo}"
        # the reference's live stepper on the same decomposition (compile-time row widths only): the single-domain bits
        if int(m["nEdgesOnCell"].max()) <= 7:
            fu, fh, fs = _run_emulated_fe(backend, md, (ssh, u, h), nparts, dt, nsteps, split_parts=bool(rng.integers(0, 2)))
            if not (np.array_equal(fu, pfe.normalVelocity) and np.array_equal(fh, pfe.layerThickness) and np.array_equal(fs, pfe.ssh)):
                problems.append(f"decomposed ForwardEuler {nparts} ranks")
        # and, now and then, the product's DecomposedModel itself: one host thread per rank, two streams per rank, in-stream
        # exchange, a random sequence of step() calls (graph replays from both parities), against the same single-domain bits
        if rng.integers(0, 2) == 0:
            calls = [int(x) for x in rng.integers(1, 5, size=int(rng.integers(1, 4)))]
            overlap, graph = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
            if rng.integers(0, 3) == 0:
                halo = "p2p_ll"                                   # the flag-in-data exchange exists in the library's own schedule only
            fe = halo in ("nccl", "p2p_ll") and int(m["nEdgesOnCell"].max()) <= 7 and bool(rng.integers(0, 2))
            step_type = mb.ForwardEuler if fe else mb.RungeKutta4
            trace("DecomposedModel", calls, halo, overlap, graph, step_type.__name__)
            locs = partition.decompose(md, nparts)

            def body(r, comm):
                model = multi_gpu.DecomposedModel(locs[r], multi_gpu.local_state(locs[r], ssh, u, h), mb.B200(0), 0, overlap=overlap,
                                                  graph=graph, runtime=simcuda.SimRuntime(comm, r), halo=halo)
                for n in calls:
                    model.step(dt, n, stepper=step_type)
                model.finish()
                res = (np.array(model.owned("normalVelocity")), np.array(model.owned("layerThickness")))
                model.close()
                return res

            outs = simcuda.run_ranks(nparts, body)
            gu2, gh2 = np.full(m["nEdges"], np.nan), np.full(m["nCells"], np.nan)
            for loc, (ru, rh) in zip(locs, outs):
                gu2[loc["edgesGlobal"][:loc["nEdgesOwned"]]] = ru
                gh2[loc["cellsGlobal"][:loc["nCellsOwned"]]] = rh
            ref = mb.PrognosticVars(ssh, u, h, 2, mb.Mesh(m, backend))
            mb.ocn_timestep(dt, ref, None, None, None, step_type, nsteps=sum(calls))
            if not (np.array_equal(gu2, ref.normalVelocity) and np.array_equal(gh2, ref.layerThickness)):
                problems.append(f"DecomposedModel {nparts} ranks {halo} overlap={overlap} graph={graph} calls={calls} {step_type.__name__}")
            desc += f" + threads {calls} {halo} overlap={overlap} graph={graph}{' ForwardEuler' if fe else ''}"
            # the reverse mode on the same decomposition (either stepper; ForwardEuler where its exchange exists)
            if rng.integers(0, 2) == 0:
                na = min(nsteps, 3)
                trace("reverse mode on the decomposition", na)

                def body_adj(r, comm):
                    model = multi_gpu.DecomposedModel(locs[r], multi_gpu.local_state(locs[r], ssh, u, h), mb.B200(0), 0, overlap=overlap,
                                                      graph=graph, runtime=simcuda.SimRuntime(comm, r), halo=halo)
                    J = model.reverse_run_loop(dt, na, stepper=step_type)
                    model.finish()
                    g = tuple(np.array(a) for a in model.gradient()) + (np.array(model.gradient_ssh()),)
                    model.close()
                    return J, g

                outs = simcuda.run_ranks(nparts, body_adj)
                au, ah, as_ = np.full(m["nEdges"], np.nan), np.full(m["nCells"], np.nan), np.full(m["nCells"], np.nan)
                for loc, (_, (ru, rh, rs)) in zip(locs, outs):
                    au[loc["edgesGlobal"][:loc["nEdgesOwned"]]] = ru
                    ah[loc["cellsGlobal"][:loc["nCellsOwned"]]] = rh
                    as_[loc["cellsGlobal"][:loc["nCellsOwned"]]] = rs
                if fe:
                    Jo, ou, oh, os_, _ = A.gradient_sum_ssh2_fe(mo, ssh, u, h, dt, na)
                    okA = rel(as_, os_) <= 1e-11
                else:
                    Jo, ou, oh = A.gradient_sum_ssh2(mo, u, h, dt, na)
                    okA = True
                okA = okA and rel(au, ou) <= 1e-11 and rel(ah, oh) <= 1e-11 and all(abs(o[0] - Jo) <= 1e-11 * abs(Jo) for o in outs)
                if not okA:
                    problems.append(f"decomposed reverse mode {nparts} ranks {halo} {step_type.__name__}")
                desc += " + reverse mode"
        # multi-level columns on the same decomposition (uniform f or not; compile-time or run-time row widths)
        if rng.integers(0, 3) == 0:
            import moka_oracle as O
            K = int(rng.integers(2, 6))
            trace("multi-level columns on the decomposition", K)
            frac = rng.uniform(0.5, 1.5, K)
            frac /= frac.sum()
            mk = dict(md)
            OC.sign_index_fields(mk)
            H = h - ssh
            rest = np.outer(H, frac)
            hk, uk = rest + np.outer(ssh, frac), np.outer(u, 1.0 + 0.1 * np.arange(K))
            mk["restingThickness"], mk["nVertLevels"] = rest, K
            locs_k = partition.decompose(mk, nparts)
            graph_k = bool(rng.integers(0, 2))
            nk = int(rng.integers(1, 4))

            def body_k(r, comm):
                model = multi_gpu.DecomposedModel(locs_k[r], multi_gpu.local_state(locs_k[r], ssh, uk, hk), mb.B200(0), 0, graph=graph_k,
                                                  runtime=simcuda.SimRuntime(comm, r), halo=str(rng_halo))
                model.step(dt, nk)
                model.finish()
                res = tuple(np.array(model.owned(f)) for f in ("normalVelocity", "layerThickness", "ssh"))
                model.close()
                return res

            rng_halo = rng.choice(["nccl", "p2p", "p2p_ll"])
            outs = simcuda.run_ranks(nparts, body_k)
            ku, kh, ks = np.full((m["nEdges"], K), np.nan), np.full((m["nCells"], K), np.nan), np.full(m["nCells"], np.nan)
            for loc, (ru, rh, rs) in zip(locs_k, outs):
                ku[loc["edgesGlobal"][:loc["nEdgesOwned"]]] = ru
                kh[loc["cellsGlobal"][:loc["nCellsOwned"]]] = rh
                ks[loc["cellsGlobal"][:loc["nCellsOwned"]]] = rs
            # against the single-domain library run (itself bit-identical to the oracle with a level axis where weights are unfolded)
            mesh_k = mb.Mesh({**m, "restingThickness": rest, "nVertLevels": K}, backend)
            ref = mb.PrognosticVars(ssh, uk, hk, 2, mesh_k)
            mb.ocn_timestep(dt, ref, None, None, None, mb.RungeKutta4, nsteps=nk)
            if not (np.array_equal(ku, np.asarray(ref.normalVelocity).reshape(m["nEdges"], K)) and
                    np.array_equal(kh, np.asarray(ref.layerThickness).reshape(m["nCells"], K)) and np.array_equal(ks, ref.ssh)):
                problems.append(f"multi-level ({K}) decomposed {nparts} ranks {rng_halo} graph={graph_k}")
            desc += f" + {K} levels decomposed ({rng_halo})"
    return f"{desc}; {nsteps} steps; renumber={renumber}; {policy}", problems


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iterations", type=int, default=30)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    backend = mb.B200(0)
    bad = 0
    for it in range(args.iterations):
        t0 = time.time()
        try:
            desc, problems = one_case(rng, backend)
        except Exception:                                      # noqa: BLE001
            desc, problems = "exception", [traceback.format_exc(limit=4)]
        bad += bool(problems)
        print(f"[{it}] {desc}: {'OK' if not problems else 'FAILED ' + '; '.join(problems)} ({time.time() - t0:.1f}s)", flush=True)
    print("FUZZ_OK" if bad == 0 else f"FUZZ_FAILED ({bad})", flush=True)
    sys.exit(0 if bad == 0 else 1)


if __name__ == "__main__":
    main()
