"""The drop-in boundary at source level: the reference's own main.cpp (AnisotropicElastoplasticity/main.cpp, unmodified,
compiled from where it lies) builds and links against the B200 host classes through include/aep/compat -- the reference's
header names forwarding to include/aep/*.h, <Eigen/Core> forwarding to the dense-type shim, and a headless igl viewer whose
launch() presses main.cpp's 's' key and pumps its pre_draw callback.  On a GPU it then runs main.cpp's scene (a 565-vertex
cloth with two pinned vertices over a ground plane, main.cpp:51-91) and writes mesh/mesh_N.obj like the reference.
The binary is built by __graft_entry__.build_reference_main() where /root/reference exists and travels to the GPU box."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

EXE = os.path.join(ROOT, "tests", "_bin", "ref_main")
REF_MAIN = "/root/reference/AnisotropicElastoplasticity/main.cpp"
have = pytest.mark.skipif(not (os.path.exists(EXE) or os.path.exists(REF_MAIN)), reason="tests/_bin/ref_main not built and /root/reference absent")
# the two bindings of INTEGRATION.md: A = main.cpp against the host classes of libaep_host.so (include/aep/compat);
# B = main.cpp + the reference's own container classes + integration/HybridSolver_b200.cpp in place of HybridSolver.cpp
BINDINGS = pytest.mark.parametrize("binding", ["A_host_classes", "B_reference_classes_patched_solve"])


def _exe(binding="A_host_classes"):
    import __graft_entry__ as g
    g.build_host()
    exe = EXE if binding.startswith("A") else EXE + "_patched"
    assert os.path.exists(exe)
    return exe


def write_square_obj(path, n=24, drop=11, side=1.0, z=0.0):
    """A triangulated square sheet with n*n - drop = 565 vertices (main.cpp:51 `nMeshParticle = 565`; the reference's
    square_hr2x06.obj is not part of its repository).  Vertices 0 and 1 -- the ones main.cpp pins -- are neighbours on an edge."""
    keep = np.ones(n * n, bool); keep[n * n - drop:] = False
    new_id = np.cumsum(keep) - 1
    xs = np.linspace(-0.5 * side, 0.5 * side, n)
    with open(path, "w") as f:
        f.write("# square sheet for main.cpp\n")
        for i in range(n):
            for j in range(n):
                if keep[i * n + j]:
                    f.write(f"v {xs[j]:.6f} {xs[i]:.6f} {z:.6f}\n")
        nf = 0
        for i in range(n - 1):
            for j in range(n - 1):
                a, b, c, d = i * n + j, i * n + j + 1, (i + 1) * n + j + 1, (i + 1) * n + j
                if keep[[a, b, c, d]].all():
                    f.write(f"f {new_id[a] + 1} {new_id[b] + 1} {new_id[c] + 1}\nf {new_id[a] + 1} {new_id[c] + 1} {new_id[d] + 1}\n"); nf += 2
    return int(keep.sum()), nf


def read_obj(path):
    v = []; f = 0
    for ln in open(path):
        if ln.startswith("v "):
            v.append([float(t) for t in ln.split()[1:4]])
        elif ln.startswith("f "):
            f += 1
    return np.array(v), f


@have
def test_reference_main_builds_against_host_library():
    exe = _exe()
    out = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True).stdout
    for sym in ("HybridSolver", "LagrangianMesh", "ParticleSystem", "RegularGrid", "groundLevelSet"):       # bound to libaep_host.so
        assert sym in out, sym
    needed = subprocess.run(["readelf", "-d", exe], capture_output=True, text=True).stdout
    assert "libaep_host.so" in needed                                   # which in turn needs libaep_b200.so (the C ABI)


@have
def test_reference_classes_build_with_the_patched_solve():
    """Binding B: everything is the reference's except HybridSolver.cpp; the loop's stage methods are gone from the binary, the
    containers are the reference's own (defined inside it), and the only library it needs is the C ABI."""
    exe = _exe("B")
    syms = subprocess.run(["nm", "-C", exe], capture_output=True, text=True).stdout
    for gone in ("evaluateInterpolationWeights_", "particleToGrid_", "computeGridForces_", "updatePlasticity_"):
        assert gone not in syms, gone
    for mine in (" T HybridSolver::solve(", " T LagrangianMesh::ObjMesh(", " T ParticleSystem::SandCylinder(", " T RegularGrid::RegularGrid("):
        assert mine in syms, mine
    needed = subprocess.run(["readelf", "-d", exe], capture_output=True, text=True).stdout
    assert "libaep_b200.so" in needed and "libaep_host.so" not in needed


@have
@BINDINGS
def test_reference_main_fails_loudly_without_gpu(tmp_path, binding):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    nv, nf = write_square_obj(str(tmp_path / "square_hr2x06.obj"))
    assert nv == 565
    r = subprocess.run([_exe(binding)], cwd=str(tmp_path), capture_output=True, text=True, timeout=120, env=dict(os.environ, AEP_HEADLESS_SECONDS="3"))
    assert r.returncode != 0 and "no CPU fallback" in r.stderr          # the solve thread dies on aep_create: nothing is simulated on the CPU
    assert not os.path.exists(tmp_path / "mesh" / "mesh_0.obj")


def main_scene(obj_path):
    """main.cpp's scene (grid main.cpp:53-69, mesh parameters :77-78, pins :71-75, ground :88-89).  The oracle's positions when
    the first 1/60 s frame completes are the expectation below: from rest the dt rule sits on its 1e-3 ceiling, so this frame
    is not chaotic."""
    import math
    from anisotropicelastoplasticity_b200 import scenes as sc
    V, F = [], []
    for ln in open(obj_path):
        if ln.startswith("v "):
            V.append([float(np.float32(t)) for t in ln.split()[1:4]])          # the loader parses positions as float (LagrangianMesh.cpp:217,229)
        elif ln.startswith("f "):
            F.append([int(t) - 1 for t in ln.split()[1:4]])
    V = np.array(V); F = np.array(F, np.int32)
    gl = 1.0 / (math.sqrt(565) - 1); mn = np.array([-2.5, -1.25, -1.67]); mx = np.array([1.25, 1.25, 1.67])
    res = np.array([int((mx[a] - mn[a]) / gl + 0.5) for a in range(3)])
    mesh = sc.make_cloth(2, 2, (0, 0, 0), (1, 0, 0), (0, 1, 0))                 # only for the material rules; geometry replaced below
    v1, v2, v3 = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
    la, lb, lc = (np.linalg.norm(a, axis=1) for a in (v2 - v1, v3 - v2, v1 - v3)); sp = 0.5 * (la + lb + lc)
    evol = 0.25 * np.sqrt(np.maximum(sp * (sp - la) * (sp - lb) * (sp - lc), 0.0)) * 0.04      # LagrangianMesh.cpp:311-315
    vvol = np.zeros(len(V))
    for c in range(3):
        np.add.at(vvol, F[:, c], evol)
    nrm = np.cross(v2 - v1, v3 - v1); nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    eD = np.stack([v2 - v1, v3 - v1, nrm]); nv, nf = len(V), len(F)
    fixed = np.zeros(nv); fixed[:2] = 1.0
    m = sc.Mesh(vx=V, vv=np.zeros((nv, 3)), vm=2e3 * vvol, vvol=vvol, vB=np.zeros((nv, 3, 3)), faces=F, ev=np.zeros((nf, 3)), em=2e3 * evol, evol=evol,
                eB=np.zeros((nf, 3, 3)), ed=eD.copy(), eD=eD, fixed=fixed, mu=mesh.mu, lam=mesh.lam, shear=0.0, stiff=4e4, fric=0.0)
    return sc.Scene("main_cpp", sc.GridSpec(mn, mx, res), sc.SAND, None, m, sc.LevelSetSpec(sc.LS_GROUND, np.array([-1.4, 0, 0, 0, 0, 0, 0, 0.0])))


def oracle_first_frame(obj_path):
    from oracle.oracle_py import Oracle
    o = Oracle(main_scene(obj_path), threads=0); o.init()
    while o.frame < 1:
        o.substep()
    return o.mesh()["vx"]


def test_oracle_first_frame_of_main_scene(tmp_path):
    """CPU: the expectation the GPU test below holds main.cpp's run to."""
    nv, nf = write_square_obj(str(tmp_path / "square_hr2x06.obj"))
    x = oracle_first_frame(str(tmp_path / "square_hr2x06.obj"))
    assert x.shape == (565, 3) and np.isfinite(x).all()
    assert x[48:, 2].mean() == pytest.approx(-0.5 * 9.8 / 3600.0, rel=0.05)             # free fall for 1/60 s, minus what the pins hold back
    assert np.abs(x[:2, 2]).max() < 1e-12                                               # pinned


def test_reference_solve_on_main_scene_matches_oracle(tmp_path):
    """CPU, live: HybridSolver::solve ITSELF (the reference's code, oracle/_ref) on main.cpp's scene for one frame, mesh_0.obj
    included, against the oracle."""
    from oracle import ref_py
    if not ref_py.available():
        pytest.skip("oracle/_ref/libaep_ref.so not built and /root/reference absent")
    write_square_obj(str(tmp_path / "square_hr2x06.obj"))
    scene = main_scene(str(tmp_path / "square_hr2x06.obj"))
    r = ref_py.Reference(scene); r.solve(0.0, str(tmp_path))
    x = oracle_first_frame(str(tmp_path / "square_hr2x06.obj"))
    assert np.abs(r.mesh()["vx"] - x).max() < 1e-10
    v, nf = read_obj(str(tmp_path / "mesh" / "mesh_0.obj"))
    assert nf == scene.mesh.nf and np.allclose(v, x, rtol=2e-5, atol=1e-9) and not os.path.exists(tmp_path / "mesh" / "mesh_1.obj")


@have
@pytest.mark.gpu
@BINDINGS
def test_reference_main_runs_on_the_engine(tmp_path, binding):
    nv, nf = write_square_obj(str(tmp_path / "square_hr2x06.obj"))
    r = subprocess.run([_exe(binding)], cwd=str(tmp_path), capture_output=True, text=True, timeout=300, env=dict(os.environ, AEP_HEADLESS_SECONDS="10"))
    assert r.returncode == 0, r.stderr[-2000:]
    assert "headless viewer" in r.stderr
    frames = sorted(os.listdir(tmp_path / "mesh"), key=lambda s: int(s.split("_")[1].split(".")[0]))
    assert len(frames) >= 4 and frames[0] == "mesh_0.obj" and not os.path.exists(tmp_path / "particle" / "particle_0.obj")   # main.cpp binds no ParticleSystem
    later = frames[min(len(frames) - 2, 15)]                              # <= 0.27 s of simulated time; the newest file may be mid-write
    v0, f0 = read_obj(str(tmp_path / "mesh" / frames[0])); v1, f1 = read_obj(str(tmp_path / "mesh" / later))
    assert v0.shape == (nv, 3) and f0 == nf and v1.shape == (nv, 3) and f1 == nf
    assert np.isfinite(v0).all() and np.isfinite(v1).all()
    assert np.abs(v0 - oracle_first_frame(str(tmp_path / "square_hr2x06.obj"))).max() < 1e-4      # frame 0 is the oracle's frame 0 (OBJ text: 6 digits)
    free = np.arange(nv) >= 48                                            # rows away from the two pinned vertices
    assert v1[free, 2].mean() < v0[free, 2].mean() < 0.0                  # the sheet falls (gravity is -z, HS:457), frame after frame
    assert np.abs(v1[:2] - np.array([[-0.5, -0.5, 0.0], [-0.5 + 1.0 / 23, -0.5, 0.0]])).max() < 0.02      # pinned vertices stay (main.cpp:71-75, HS:513-550)
