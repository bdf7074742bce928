"""The C ABI is usable from plain C: tests/cabi_driver.c is compiled with gcc (-std=c99 -pedantic -Werror) against
include/gkr_msm_b200.h, linked with the in-tree library, and -- on the GPU -- replays BareSumcheckSO::prove
(src/cleanup/protocols/sumcheck.rs:646-691 over GenericSumcheckProtocol::prove, :101-123) through the trait-shaped entries
the way the reference's Rust host would (unipoly -> from_evals -> compress -> write_scalars -> challenge -> bind).  Its proof
bytes must equal the python oracle's."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cabi_driver.c")
BIN = os.path.join(ROOT, "tests", "cabi_driver")
LIBDIR = os.path.join(ROOT, "gkr-msm_b200", "lib")


def build_driver(out=BIN):
    subprocess.check_call(["gcc", "-std=c99", "-O1", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", out,
                           "-L", LIBDIR, "-lgkr_msm_b200", "-Wl,-rpath," + LIBDIR])
    return out


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_header_compiles_as_c99_and_links(tmp_path):
    """no compute call: the header is valid ISO C99, and every entry the driver uses resolves against the library"""
    assert os.path.exists(os.path.join(LIBDIR, "libgkr_msm_b200.so")), "build the library first (__graft_entry__.build())"
    out = build_driver(str(tmp_path / "cabi_driver"))
    undefined = subprocess.run(["nm", "-u", out], capture_output=True, text=True).stdout
    used = sorted({line.split()[-1] for line in undefined.splitlines() if " gkr_" in line or line.strip().startswith("U gkr_")})
    assert {"gkr_so_unipoly", "gkr_so_bind", "gkr_so_final_evals", "gkr_transcript_challenge", "gkr_sumcheck_prove"} <= {u.split("@")[0] for u in used}
    exported = subprocess.run(["nm", "-D", "--defined-only", os.path.join(LIBDIR, "libgkr_msm_b200.so")], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in exported.splitlines()}
    for sym in used:
        assert sym.split("@")[0] in exported, sym


@pytest.mark.gpu
@pytest.mark.parametrize("nv,seeds", [(5, (11, 12, 13)), (13, (7, 8, 9))])  # 13: rounds above and below the small-round kernel threshold
def test_c_driver_proof_equals_the_oracle(nv, seeds):
    from oracle.pyref import gates as G
    from oracle.pyref import sumcheck as S
    from oracle import coracle
    from oracle.pyref.field import P, fr_vec_from_mont_u64
    from oracle.pyref.transcript import ProofTranscript2

    exe = BIN if os.path.exists(BIN) else build_driver()
    res = subprocess.run([exe, str(nv)] + [str(s) for s in seeds], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    got = bytes.fromhex(res.stdout.strip())
    n = 1 << nv
    tabs = [fr_vec_from_mont_u64(coracle.synth_table(s, n)) for s in seeds]  # the generator shared with gkr_table_synth
    claim = sum(a * b % P * c for a, b, c in zip(*tabs)) % P
    tr = ProofTranscript2.start_prover(b"fgstglsp")
    so = S.DenseSumcheckObjectSO(tabs, G.Prod3(), nv, claim)
    S.BareSumcheckSO(G.Prod3(), nv).prove(tr, claim, so)
    assert got == tr.end()


def _c_decls():
    import re
    src = open(os.path.join(ROOT, "include", "gkr_msm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(gkr_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return out


def test_rust_ffi_declarations_match_the_header():
    """rust/src/gpu/ffi.rs cannot be compiled in this image (no cargo): at least every `extern "C"` item must name an entry
    the header declares, with the same number of arguments."""
    import re
    c = _c_decls()
    rs = open(os.path.join(ROOT, "rust", "src", "gpu", "ffi.rs")).read()
    rs = re.sub(r"//[^\n]*", "", rs)
    found = 0
    for m in re.finditer(r"pub fn (gkr_\w+)\s*\(([^)]*)\)", rs, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        n = 0 if not args else len([a for a in args.split(",") if a.strip()])
        assert name in c, f"{name} is not declared in include/gkr_msm_b200.h"
        assert c[name] == n, f"{name}: {n} arguments in ffi.rs, {c[name]} in the header"
        found += 1
    assert found >= 25
