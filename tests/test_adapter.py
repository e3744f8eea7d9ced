"""The PETSc-facing boundary of the reference (hdr/geneo_c.h, hdr/geneo.hpp) re-exported by
geneo4petsc_b200/csrc/petsc_adapter.cpp, compiled against the reference's own headers and the PETSc/MPI stand-in of
tests/petsc_stub (this image has neither PETSc nor MPI).  CPU: symbols, registration, option grammar, defaults, usage.
GPU: P thread-ranks drive PCSetUp / PCApply through pc->ops and must reproduce geneo_pc_apply on the same decomposition."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADAPTER = os.path.join(ROOT, "geneo4petsc_b200", "libgeneo_petsc_adapter.so")
DRIVER = os.path.join(ROOT, "build", "adapter_driver")

needs_build = pytest.mark.skipif(not (os.path.exists(ADAPTER) and os.path.exists(DRIVER)),
                                 reason="adapter not built (needs the reference headers at build time: /root/reference/hdr)")


@needs_build
def test_adapter_exports_the_reference_symbols():
    """createGenEOPC / PCGenEOSetup with C linkage (hdr/geneo_c.h:9-10), initGenEOPC / usageGenEO with the C++ mangling of
    hdr/geneo.hpp:30-41 (PETSc's own struct names _p_PC, _p_Mat, _p_Vec, _p_ISLocalToGlobalMapping)."""
    out = subprocess.run(["nm", "-D", "--defined-only", ADAPTER], capture_output=True, text=True, check=True).stdout
    syms = {ln.split()[-1] for ln in out.splitlines() if ln.strip()}
    assert "createGenEOPC" in syms and "PCGenEOSetup" in syms
    init = [s for s in syms if s.startswith("_Z11initGenEOPC")]
    assert len(init) == 1 and "_p_PC" in init[0] and "_p_ISLocalToGlobalMapping" in init[0] and "_p_Mat" in init[0] and "_p_Vec" in init[0]
    assert any(s.startswith("_Z10usageGenEO") for s in syms)


@needs_build
def test_adapter_registration_options_and_context_on_cpu():
    r = subprocess.run([DRIVER, "cpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "adapter cpu ok" in r.stdout, (r.stdout, r.stderr)


@needs_build
@pytest.mark.gpu
@pytest.mark.parametrize("args", [("4", "10", "ASM,1", "0", "0"), ("4", "10", "SORAS,2", "1", "0"), ("3", "9", "ASM,H1", "0", "1"),
                                  ("2", "12", "RAS,0", "1", "1")])
def test_adapter_setup_and_apply_through_pc_ops(args):
    r = subprocess.run([DRIVER, "gpu"] + list(args), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-800:], r.stderr[-800:])
    assert "failures 0" in r.stdout
