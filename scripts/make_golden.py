#!/usr/bin/env python
"""Generate ``tests/golden/*`` from the reference tree (run in the build container).

Reads ONLY data artefacts of the reference (no source is copied):

* ``/root/reference/data/simdata.rda``            -> ``simdata.npz``
* ``/root/reference/data/yeast_benchmark.rda``    -> ``yeast_benchmark.npz``
* ``/root/reference/tests/testdata/ref_bces.rda`` -> ``ref_bces.json``
  (W, H, thresholds, memberships of the als-nmf / svd-pca / nipals-pca / snmf
  strategies the reference's own testthat file compares against -- for snmf
  also the NMFfit's start seed, parameters, iteration count and objective --
  plus the 6x6 input matrix)
* ``/root/reference/vignettes/my-vignette.html``  -> ``vignette.json``
  (printed outputs of ``addStrat(k=3, 'als-nmf')`` and ``auto_bcv`` under
  ``set.seed(12345)`` on ``yeast_benchmark[[1]]``)

``/root/reference`` does not exist on the GPU box; the committed outputs do.
"""
import html
import json
import os
import re
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle.rdx import load_rda  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    sim = load_rda(f"{REF}/data/simdata.rda")["simdata"]
    np.savez_compressed(f"{OUT}/simdata.npz",
                        **{n: v.array() for n, v in zip(sim.names(), sim.value)})

    yeast = load_rda(f"{REF}/data/yeast_benchmark.rda")["yeast_benchmark"]
    arrs = {}
    for i, v in enumerate(yeast.value):
        arrs[f"y{i + 1:02d}"] = v.array()
    dn = yeast.value[0].attr["dimnames"].value
    arrs["y01_rownames"] = np.array(dn[0].value)
    arrs["y01_colnames"] = np.array(dn[1].value)
    np.savez_compressed(f"{OUT}/yeast_benchmark.npz", **arrs)

    ref = load_rda(f"{REF}/tests/testdata/ref_bces.rda")["ref_bces"]
    out = {}
    for key in ("bce.als_nmf", "bce.svd_pca", "bce.nipals_pca"):
        bce = ref[key]
        strat = bce.slot("strategies").value[0]
        fit = strat.slot("factors").slot("fit")
        bc = strat.slot("biclust")
        out[key] = dict(
            name=bce.slot("strategies").names()[0],
            W=fit.slot("W").array().tolist(),
            H=fit.slot("H").array().tolist(),
            featureThresh=np.asarray(strat.slot("featureThresh").value).tolist(),
            sampleThresh=np.asarray(strat.slot("sampleThresh").value).tolist(),
            RowxNumber=bc.slot("RowxNumber").array().astype(int).tolist(),
            NumberxCol=bc.slot("NumberxCol").array().astype(int).tolist(),
            k=int(strat.slot("k").value[0]),
        )
    # bce.snmf: the stored NMFfit of NMF::nmf(method = "snmf/r") -- start seed, parameters, factors, iteration count
    # and final objective (test_BiclusterExperiment.R:16, :43-54)
    bce = ref["bce.snmf"]
    strat = bce.slot("strategies").value[0]
    nf = strat.slot("factors")
    fit = nf.slot("fit")
    bc = strat.slot("biclust")
    par = nf.slot("parameters")
    out["bce.snmf"] = dict(
        name=bce.slot("strategies").names()[0],
        W=fit.slot("W").array().tolist(),
        H=fit.slot("H").array().tolist(),
        beta=float(np.asarray(par.value[par.names().index("beta")].value)[0]),
        eta=float(np.asarray(par.value[par.names().index("eta")].value)[0]),
        iterations=int(np.asarray(nf.slot("extra").value[0].value)[0]),
        residual=float(np.asarray(nf.slot("residuals").value)[-1]),
        random_seed=[int(v) for v in np.asarray(nf.slot("rng").value)],
        featureThresh=np.asarray(strat.slot("featureThresh").value).tolist(),
        sampleThresh=np.asarray(strat.slot("sampleThresh").value).tolist(),
        RowxNumber=bc.slot("RowxNumber").array().astype(int).tolist(),
        NumberxCol=bc.slot("NumberxCol").array().astype(int).tolist(),
        k=int(strat.slot("k").value[0]),
    )
    abund = ref["bce.als_nmf"].slot("assayData")
    # assayData is a list (storageMode "list") or environment holding `abund`
    a = abund["abund"] if abund.kind == "list" else abund.value["abund"]
    out["input"] = a.array().tolist()
    with open(f"{OUT}/ref_bces.json", "w") as fh:
        json.dump(out, fh, indent=1)

    # ---- vignette -----------------------------------------------------------
    raw = open(f"{REF}/vignettes/my-vignette.html", encoding="utf-8").read()
    txt = html.unescape(re.sub(r"<[^>]*>", "", raw))
    lines = [ln[3:].rstrip() if ln.startswith("#> ") else ln.rstrip() for ln in txt.splitlines()]
    traces, cur = [], []
    for ln in lines:
        mt = re.match(r"^\s+(\d+)\s+(\d+\.\d+)\s+(\d+\.\d+)\s+(\d+\.\d+)$", ln)
        if mt:
            cur.append([int(mt.group(1))] + [float(mt.group(i)) for i in (2, 3, 4)])
        elif ln.startswith("Converged!"):
            traces.append(cur)
            cur = []

    def table(start_pat, row_prefix, conv):
        """collect `Bicluster.N v v v` style blocks following a line matching start_pat"""
        rows = {}
        on = False
        for ln in lines:
            if re.search(start_pat, ln):
                on = True
                continue
            if on:
                if ln.startswith(row_prefix):
                    parts = ln.split()
                    rows.setdefault(parts[0], []).extend(conv(p) for p in parts[1:])
                elif ln.strip() == "" or not (ln.startswith(" ") or ln.startswith(row_prefix)):
                    if rows and not ln.startswith(" "):
                        break
        return rows

    sf = table(r"^sampleFactor\(bcs\)", "Bicluster.", float)
    cs = table(r"^clusteredSamples\(bcs\)", "Bicluster.", lambda s: s == "TRUE")

    def head_rows(start_pat, conv):
        rows = []
        on = False
        for ln in lines:
            if re.search(start_pat, ln):
                on = True
                continue
            if on:
                if re.match(r"^Y[A-Z0-9\-]+\s", ln):
                    parts = ln.split()
                    rows.append([parts[0]] + [conv(p) for p in parts[1:]])
                elif "reached getOption" in ln:
                    break
        return rows

    ff = head_rows(r"^featureFactor\(bcs\)", float)
    cf = head_rows(r"^clusteredFeatures\(bcs\)", lambda s: s == "TRUE")

    def name_list(start_pat):
        names, omitted, on = [], 0, False
        for ln in lines:
            if re.search(start_pat, ln):
                on = True
                continue
            if on:
                mo = re.search(r"omitted (\d+) entries", ln)
                if mo:
                    omitted = int(mo.group(1))
                    break
                found = re.findall(r'"([^"]+)"', ln)
                if not found and names:
                    break
                names.extend(found)
        return names, omitted

    b1_names, b1_om = name_list(r"^names\(which\(clusteredFeatures\(bcs\)\[, 1\]\)\)")
    b2_names, b2_om = name_list(r"^names\(which\(clusteredFeatures\(bcs\)\[, 2\]\)\)")
    best = counts = None
    for i, ln in enumerate(lines):
        if ln.strip() == "$best":
            best = int(lines[i + 1].split()[1])
        if ln.strip() == "$counts":
            counts = dict(zip([int(x) for x in lines[i + 1].split()],
                              [int(x) for x in lines[i + 2].split()]))
    vig = dict(seed=12345, shift=2.71, k=3, reps=4, traces=traces,
               sampleFactor={k: v for k, v in sf.items()},
               featureFactor_head=ff,
               clusteredSamples={k: v for k, v in cs.items()},
               clusteredFeatures_head=cf,
               b1_feature_names_head=b1_names, b1_feature_count=len(b1_names) + b1_om,
               b2_feature_names_head=b2_names, b2_feature_count=len(b2_names) + b2_om,
               auto_bcv_best=best, auto_bcv_counts=counts)
    with open(f"{OUT}/vignette.json", "w") as fh:
        json.dump(vig, fh, indent=1)
    print("traces", [len(t) for t in traces], "sf", {k: len(v) for k, v in sf.items()},
          "ff", len(ff), "cs", {k: len(v) for k, v in cs.items()}, "cf", len(cf),
          "b1", len(b1_names), b1_om, "b2", len(b2_names), b2_om, best, counts)


if __name__ == "__main__":
    main()
