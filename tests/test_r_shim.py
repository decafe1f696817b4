"""The R side of the drop-in boundary exists as files (r/src/r_shim.c, r/R/backends.R, r/src/Makevars) and is kept
consistent without an R installation: the shim is type-checked against include/mfbiclust.h with minimal stand-in
declarations of R's C API (r/stub/), and every .Call target of the R code must be registered in the shim with the
number of arguments the R code passes."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_r_shim_type_checks_against_the_c_abi_header():
    cmd = ["gcc", "-fsyntax-only", "-Wall", "-Werror", "-std=c11", "-I" + os.path.join(ROOT, "r", "stub"),
           "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "r", "src", "r_shim.c")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def _split_args(s):
    depth, cur, out = 0, "", []
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    out.append(cur)
    return out


def test_r_calls_match_the_registration_table():
    shim = open(os.path.join(ROOT, "r", "src", "r_shim.c")).read()
    table = {m.group(1): int(m.group(2)) for m in re.finditer(r'\{"(C_mfb_\w+)", \(DL_FUNC\) &\1, (\d+)\}', shim)}
    assert len(table) >= 16
    # each registered routine is defined with that many SEXP parameters
    for name, nargs in table.items():
        m = re.search(r"SEXP %s\(([^)]*)\)" % name, shim)
        assert m, name
        assert m.group(1).count("SEXP") == nargs, name
    rsrc = open(os.path.join(ROOT, "r", "R", "backends.R")).read()
    calls = 0
    for m in re.finditer(r"\.Call\(", rsrc):
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(rsrc[i], 0)
            i += 1
        args = _split_args(rsrc[m.end():i - 1])
        name = args[0].strip()
        assert name in table, name
        assert len(args) - 1 == table[name], (name, len(args) - 1, table[name])
        calls += 1
    assert calls >= 14
    # the functions of the reference's drop-in surface are all defined (NAMESPACE:5-7,20-21,28-29,42)
    for fn in ("als_nmf", "svd_pca", "nipals_pca", "nipals_pca_nocatch", "bcv", "auto_bcv", "bcvGivenKs", "otsuHelper",
               "threshold"):
        assert re.search(r"^%s <- function\(" % fn, rsrc, re.M), fn


def test_makevars_links_the_library():
    mk = open(os.path.join(ROOT, "r", "src", "Makevars")).read()
    assert "-lmfbiclust_cuda" in mk and "include" in mk
