"""Host-side plumbing of the two bindings, checked WITHOUT a GPU against a call recorder (tests/recorder/cabi_recorder.cpp).

The recorder has the C ABI's symbol names and computes nothing -- no physics, no oracle: it logs calls, keeps what it is handed and
"downloads" recognisable patterns of the uploads.  It is built here as a stand-in libaep_b200.so in a temporary directory and put in
front of the real one with LD_LIBRARY_PATH for these tests only.  What is checked is what does not need a GPU to be wrong: the call
sequence of HybridSolver::solve (HybridSolver.cpp:827-1034), which container arrays reach the ABI in which layout, the collider
samples, how many frames a given maxt yields, what lands in the OBJ files and in the containers afterwards.
  binding A: libaep_host.so's HybridSolver (tests/host_driver.cpp, mode `solve`)
  binding B: the reference's own container classes + integration/HybridSolver_b200.cpp (tests/ref_patch_driver.cpp; needs
             /root/reference to build, like oracle/_ref)
Numerical behaviour of the real library is the business of the `-m gpu` tests; nothing here stands in for it."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden
from test_host_cpp import DRIVER, PKG, _build, read_blob, scene_blob, write_blob

REF = "/root/reference/AnisotropicElastoplasticity"
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


@pytest.fixture(scope="module")
def recorder_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("recorder")
    subprocess.check_call([CXX, "-O1", "-std=c++17", "-fPIC", "-shared", "-w", "-o", str(d / "libaep_b200.so"), os.path.join(ROOT, "tests", "recorder", "cabi_recorder.cpp")])
    return d


@pytest.fixture(scope="module")
def patched_driver(tmp_path_factory, recorder_dir):
    if not os.path.exists(os.path.join(REF, "HybridSolver.h")):
        pytest.skip("/root/reference absent: binding B cannot be compiled here")
    import __graft_entry__ as g
    g.build_host()                                                     # tests/_bin/patched_obj/*.o: the reference's classes + the replacement solve
    obj = os.path.join(ROOT, "tests", "_bin", "patched_obj")
    exe = str(tmp_path_factory.mktemp("patched") / "ref_patch_driver")
    inc = ["-I" + os.path.join(ROOT, "include", "aep", "headless"), "-I" + os.path.join(ROOT, "oracle", "ref_shim"), "-I" + os.path.join(ROOT, "include"), "-I" + REF]
    objs = [os.path.join(obj, f + ".o") for f in ("ParticleSystem", "RegularGrid", "LagrangianMesh", "geometry", "interpolation", "LevelSet", "HybridSolver_b200")]
    subprocess.check_call([CXX, "-O1", "-std=c++14", "-w"] + inc + ["-o", exe, os.path.join(ROOT, "tests", "ref_patch_driver.cpp")] + objs + ["-L" + PKG, "-laep_b200"])
    return exe


def run(binding, recorder_dir, patched_driver, scene, tmp_path, n_frames):
    work = tmp_path / binding; work.mkdir()
    sin, sout, rec = str(work / "scene.bin"), str(work / "out.bin"), str(work / "recorded.bin")
    write_blob(sin, scene_blob(scene, ls_mode=1))                      # ls_mode 1: the collider as std::function, as main.cpp binds it
    env = dict(os.environ, LD_LIBRARY_PATH=str(recorder_dir), AEP_RECORDER_OUT=rec)
    maxt = n_frames / 60.0 - 1.0 / 120.0                               # `while (t <= maxt)`, t += 1/60 per frame -> n_frames frames
    if binding == "A":
        _build()
        cmd = [DRIVER, "run", sin, sout, "solve", str(n_frames), str(work)]
    else:
        cmd = [patched_driver, sin, sout, repr(maxt)]
    r = subprocess.run(cmd, cwd=str(work), env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    recd = read_blob(rec)
    recd["calls"] = "".join(chr(int(c)) for c in recd["calls"]).split()
    return read_blob(sout), recd, work


@pytest.mark.parametrize("binding", ["A", "B"])
def test_solve_plumbing(binding, recorder_dir, tmp_path, request):
    patched = request.getfixturevalue("patched_driver") if binding == "B" else None
    d, scene = load_golden("cloth_sand")                               # particles + cloth with two pinned vertices + ground plane
    n_frames = 3
    out, rec, work = run(binding, recorder_dir, patched, scene, tmp_path, n_frames)
    p, m, g = scene.particles, scene.mesh, scene.grid
    n, nv, nf, ng = p.n, m.nv, m.nf, g.n_nodes
    cm = lambda a: np.ascontiguousarray(np.asarray(a, np.float64).T).ravel()          # (N,3) -> column-major, leading dimension N
    # ---- the call sequence of solve()
    calls = rec["calls"]
    head = ["aep_create", "aep_upload_particles", "aep_upload_mesh", "aep_set_levelset_samples", "aep_init"]
    assert calls[:5] == head and calls[-1] == "aep_destroy"
    assert calls.count("aep_run_frames") == n_frames == int(rec["cfg"][11])
    # positions only, once per frame: binding A through the pinned float32 frame download, binding B through the fp64 download
    frame_calls = (["aep_run_frames", "aep_download_mesh(x)", "aep_frame_positions_begin", "aep_frame_positions_wait"] if binding == "A"
                   else ["aep_run_frames", "aep_download_particles(x)", "aep_download_mesh(x)"])
    nfc = len(frame_calls)
    per_frame = calls[5:5 + nfc * n_frames]
    assert per_frame == frame_calls * n_frames
    assert calls[5 + nfc * n_frames:-1] == ["aep_download_particles(all)", "aep_download_mesh(all)", "aep_download_grid"]
    # ---- configuration: the grid of the RegularGrid, CFL, SAND (HS:873)
    assert rec["cfg"][0] == 1 and rec["cfg"][1] == scene.cfl and np.allclose(rec["cfg"][2:8], np.concatenate([g.mn, g.mx])) and list(rec["cfg"][8:11]) == list(g.res)
    # ---- what reached the ABI: the containers' arrays in the reference's layouts
    assert np.array_equal(rec["x"], cm(p.x)) and np.array_equal(rec["v"], cm(p.v)) and np.array_equal(rec["m"], p.m) and np.array_equal(rec["q"], p.q)
    for a in range(3):
        assert np.array_equal(rec[f"B{a + 1}"], cm(p.B[:, a, :]))
    assert np.array_equal(rec["FE"], np.asarray(p.FE).transpose(0, 2, 1).ravel()) and np.array_equal(rec["FP"], np.asarray(p.FP).transpose(0, 2, 1).ravel())
    assert np.allclose(rec["mat"], [p.E, p.nu, p.thetaC, p.thetaS])
    assert np.array_equal(rec["vx"], cm(m.vx)) and np.array_equal(rec["vv"], cm(m.vv)) and np.array_equal(rec["vm"], m.vm) and np.array_equal(rec["evol"], m.evol)
    assert np.array_equal(rec["faces"], np.ascontiguousarray(m.faces.T).ravel())                                   # nf x 3 column-major int32
    assert np.array_equal(rec["ed"], np.concatenate([cm(m.ed[a]) for a in range(3)])) and np.array_equal(rec["eD"], np.concatenate([cm(m.eD[a]) for a in range(3)]))
    assert np.array_equal(rec["vB"], np.concatenate([cm(m.vB[:, a, :]) for a in range(3)]))
    assert np.array_equal(rec["fixedv"], m.fixed) and m.fixed.sum() == 2                                           # bindConstraints -> vertexIsFixed
    assert np.allclose(rec["mpar"], [m.mu, m.lam, m.shear, m.stiff, m.fric])
    # ---- the collider, sampled at the nodes (ground plane z <= z0: phi <= 0, normal +z; elsewhere nothing)
    z0 = scene.levelset.params[0]
    k = np.arange(ng) // (g.res[0] * g.res[1])
    inside = (g.mn[2] + k * g.h[2]) - z0 <= 0.0
    assert np.array_equal(rec["inside"].astype(bool), inside)
    nrm = rec["normal"].reshape(3, ng)
    assert np.array_equal(nrm[2], inside.astype(float)) and not nrm[:2].any()
    # ---- frames: particle_N.obj / mesh_N.obj for N = 0..n_frames-1, holding what the per-frame download returned (x + 0.01 (N+1))
    assert sorted(os.listdir(work / "particle")) == [f"particle_{i}.obj" for i in range(n_frames)]
    assert sorted(os.listdir(work / "mesh")) == [f"mesh_{i}.obj" for i in range(n_frames)]
    for i in range(n_frames):
        xs = np.array([[float(t) for t in ln.split()[1:]] for ln in open(work / "particle" / f"particle_{i}.obj")])
        assert xs.shape == (n, 3) and np.allclose(xs, p.x + 0.01 * (i + 1), rtol=2e-5)                              # "v x y z", 6 significant digits (HS:1003-1005)
        lines = open(work / "mesh" / f"mesh_{i}.obj").read().splitlines()
        vs = np.array([[float(t) for t in ln.split()[1:]] for ln in lines if ln.startswith("v ")])
        fs = np.array([[int(t) for t in ln.split()[1:]] for ln in lines if ln.startswith("f ")])
        assert np.allclose(vs, m.vx + 0.02 * (i + 1), rtol=2e-5) and np.array_equal(fs, m.faces + 1)                # faces 1-based (HS:1021-1023)
        assert lines[:nv] == [ln for ln in lines if ln.startswith("v ")]                                            # vertices first, then faces
    # ---- the containers after solve(): what the final downloads returned
    un = lambda a, k_: a.reshape(3, k_).T
    assert np.allclose(un(out["x"], n), p.x + 0.01 * n_frames) and np.allclose(un(out["v"], n), -p.x)
    assert np.allclose(out["FE"], 2 * np.asarray(p.FE).transpose(0, 2, 1).ravel()) and np.allclose(out["FP"], 3 * np.asarray(p.FP).transpose(0, 2, 1).ravel())
    assert np.allclose(out["vol"], p.vol + 5) and np.allclose(out["q"], p.q + 7)
    assert np.allclose(un(out["mesh_vx"], nv), m.vx + 0.02 * n_frames) and np.allclose(un(out["mesh_vv"], nv), -m.vx)
    assert np.allclose(out["mesh_ex"], 0.5) and np.allclose(un(out["mesh_d3"], nf), 2 * m.ed[2])
    assert np.allclose(out["grid_m"], 11.0) and np.allclose(out["grid_v"], 12.0)
    if binding == "B":
        assert np.allclose(un(out["B1"], n), p.B[:, 0, :] + 1) and np.allclose(un(out["B3"], n), p.B[:, 2, :] + 3)
        assert np.allclose(un(out["mesh_vB1"], nv), m.vB[:, 0, :] + 1) and np.allclose(un(out["mesh_eB3"], nf), m.eB[:, 2, :] + 2)
        assert np.allclose(un(out["mesh_ev"], nf), m.ev + 4) and np.allclose(un(out["mesh_d1"], nf), 2 * m.ed[0]) and np.allclose(out["grid_f"], 13.0)
