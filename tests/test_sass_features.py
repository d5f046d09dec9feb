"""What the compiled library (sm_100a SASS, read with cuobjdump: no GPU needed) must contain for DESIGN.md section 3 to be true:
the two gather kernels fetch their grid boxes with TMA (UTMALDG.4D) and their particle records with cp.async (LDGSTS), the two
scatter kernels accumulate in packed fp32x2 (FFMA2) and reduce whole nodes with one vector reduction (REDG.E.ADD.F32x4), and the
default particle kernels do not spill."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "anisotropicelastoplasticity_b200", "libaep_b200.so")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB), reason="cuobjdump or the built library missing")


@pytest.fixture(scope="module")
def sass():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    out = {}
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        ops = collections.Counter()
        for line in f.splitlines():
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
            if m:
                ops[m.group(1)] += 1
        out[name] = ops
    return out


def _kernel(sass, *parts):
    hits = [k for k in sass if all(p in k for p in parts)]
    assert len(hits) == 1, (parts, hits)
    return sass[hits[0]]


def _has(ops, prefix):
    return sum(v for k, v in ops.items() if k.startswith(prefix))


def test_gather_kernels_use_tma_and_cp_async(sass):
    for parts in (("k_forcesILi8ELb0ELb1",), ("k_g2p2gILi8ELb0ELb0",)):                # force gather (SPLIT), G2P: 8 rounds, cell-sorted particles
        ops = _kernel(sass, *parts)
        assert _has(ops, "UTMALDG.4D") >= 1 and _has(ops, "SYNCS.") >= 2               # the box by TMA, completion through an mbarrier
        assert _has(ops, "LDGSTS.E.BYPASS.128") >= 7                                   # X / F_E / constants by 16-byte cp.async.cg
        assert _has(ops, "LDG.") < 16                                                  # F_P rows of yielding particles only: no 64-node global-memory gather here
    lst = _kernel(sass, "k_forcesILi1ELb1ELb1")                                        # the list pass gathers from global memory instead
    assert _has(lst, "UTMALDG") == 0 and _has(lst, "LDG.") >= 16


def test_scatter_kernels_are_packed_and_reduce_whole_nodes(sass):
    for parts in (("k_p2gILi8",), ("k_force_scatterILi8",)):
        ops = _kernel(sass, *parts)
        assert _has(ops, "FFMA2") >= 24
        assert _has(ops, "REDG.E.ADD.F32x4") >= 8                                      # one reduction per node: (m, px, py, pz) or (fx, fy, fz, -)
        assert _has(ops, "ATOMS") == 0                                                 # nothing is accumulated with shared-memory atomics (CAS loops on sm_100)


def test_default_particle_kernels_do_not_spill():
    log = os.path.join(ROOT, "anisotropicelastoplasticity_b200", "csrc", "build.log")
    if not os.path.exists(log):
        pytest.skip("no ptxas log (library built elsewhere)")
    txt = open(log).read()
    blocks = re.findall(r"Compiling entry function '([^']+)'[^\n]*\n(?:ptxas info\s*: Function properties[^\n]*\n)?\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s*: Used (\d+) registers", txt)
    seen = {}
    for name, stack, sst, sld, regs in blocks:
        seen[name] = (int(sst), int(regs))
    want = {"k_forcesILi8ELb0ELb1": 64, "k_force_scatterILi8": 64, "k_p2gILi8": 64, "k_g2p2gILi8ELb0ELb0": 96}
    for key, max_regs in want.items():
        hits = [(n, v) for n, v in seen.items() if key in n]
        assert len(hits) == 1, key
        spills, regs = hits[0][1]
        assert regs <= max_regs and spills <= 16, (key, regs, spills)
