#!/usr/bin/env python
"""Convert the reference's own test fixtures for the CI-test hot path into compact
golden files that travel with this repo (the GPU box has no /root/reference).

Sources (all data, no code), relative to /root/reference/test/data:
  preprocessing_expected/{pres_abs,clr_nonzero_binned,clr_adapt,clr_nonzero}.tsv
      -> the exact *inputs* of test/tests.jl and test/learning.jl for mi / mi_nz / fz / fz_nz
  tests_expected.tsv            -> 204 golden TestResults (test/tests.jl:41-74)
  learning_expected/*.edgelist  -> 8 expected graphs (test/learning.jl:176-237)

Run once in the build container:  python tests/golden/make_golden.py
  HMP_SRA_gut/HMP_SRA_gut_small.tsv + preprocessing_expected/*.tsv
      -> raw counts and the six expected normalisations (test/preprocessing.jl:48-84) for the normalisation step

Outputs: tests/golden/hmp_inputs.npz, tests/golden/tests_expected.json,
         tests/golden/learning_expected.json, tests/golden/edgelists/*.edgelist (verbatim), tests/golden/prep_fixtures.npz,
         tests/golden/meta_onehot.json
"""
import json
import os
import sys

import numpy as np

REF = os.environ.get("FW_REFERENCE", "/root/reference")
D = os.path.join(REF, "test", "data")
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    pe = os.path.join(D, "preprocessing_expected")
    inputs = {
        "mi": np.loadtxt(os.path.join(pe, "pres_abs.tsv"), delimiter="\t").astype(np.int8),
        "mi_nz": np.loadtxt(os.path.join(pe, "clr_nonzero_binned.tsv"), delimiter="\t").astype(np.int8),
        # text holds Float32-rounded values; keep them as float32 (exact round trip)
        "fz": np.loadtxt(os.path.join(pe, "clr_adapt.tsv"), delimiter="\t").astype(np.float32),
        "fz_nz": np.loadtxt(os.path.join(pe, "clr_nonzero.tsv"), delimiter="\t").astype(np.float32),
    }
    for k, v in inputs.items():
        assert v.shape == (346, 50), (k, v.shape)
    np.savez_compressed(os.path.join(OUT, "hmp_inputs.npz"), **inputs)

    raw = [l.rstrip("\n").split("\t") for l in open(os.path.join(D, "HMP_SRA_gut", "HMP_SRA_gut_small.tsv"))]
    counts = np.array([[float(v) for v in r[1:]] for r in raw[1:]])
    assert counts.shape == (351, 50) and (counts == np.round(counts)).all()
    prep = {"counts": counts.astype(np.int32)}
    for mode, fn in [("clr-adapt", "clr_adapt"), ("clr-nonzero", "clr_nonzero"), ("clr-nonzero-binned", "clr_nonzero_binned"),
                     ("pres-abs", "pres_abs"), ("tss", "tss"), ("tss-nonzero-binned", "tss_nonzero_binned")]:
        e = np.loadtxt(os.path.join(pe, fn + ".tsv"), delimiter="\t")
        prep[mode] = e.astype(np.int8) if ("binned" in mode or mode == "pres-abs") else e.astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "prep_fixtures.npz"), **prep)

    # meta variables: raw factor table and the reference's expected one-hot encoding (test/preprocessing.jl:144-170)
    def conv(v):
        for t in (int, float):
            try:
                return t(v)
            except ValueError:
                pass
        return v
    raw = [l.rstrip("\n").split("\t") for l in open(os.path.join(D, "HMP_SRA_gut", "HMP_SRA_gut_tiny_meta_oneHotTest.tsv"))]
    exp_oh = [l.rstrip("\n").split("\t") for l in open(os.path.join(pe, "meta_tiny_oneHotTest.tsv"))]
    with open(os.path.join(OUT, "meta_onehot.json"), "w") as f:
        tiny = [l.rstrip("\n").split("\t") for l in open(os.path.join(D, "HMP_SRA_gut", "HMP_SRA_gut_tiny.tsv"))]
        json.dump({"counts": [[int(v) for v in r] for r in tiny[1:]],
                   "header": raw[0], "columns": [[conv(r[j]) for r in raw[1:]] for j in range(len(raw[0]))],
                   "expected_header": exp_oh[0], "expected": [[float(v) for v in r] for r in exp_oh[1:]]}, f)

    exp = {}
    with open(os.path.join(D, "tests_expected.tsv")) as f:
        header = f.readline().rstrip("\n").split("\t")
        assert header == ["key", "stat", "pval", "df", "suff_power"], header
        for line in f:
            key, stat, pval, df, sp = line.rstrip("\n").split("\t")
            exp.setdefault(key, []).append([float(stat), float(pval), int(float(df)), sp.strip() == "true"])
    assert sum(len(v) for v in exp.values()) == 204
    with open(os.path.join(OUT, "tests_expected.json"), "w") as f:
        json.dump(exp, f, indent=0)

    graphs = {}
    ld = os.path.join(D, "learning_expected")
    for fn in sorted(os.listdir(ld)):
        name = os.path.splitext(fn)[0]
        edges = []
        with open(os.path.join(ld, fn)) as f:
            for line in f:
                if line.startswith("#") or not line.strip():
                    continue
                a, b, w = line.rstrip("\n").split("\t")
                a, b = int(a[1:]) - 1, int(b[1:]) - 1      # "X17" -> 0-based 16
                edges.append([min(a, b), max(a, b), float(w)])
        graphs[name] = sorted(edges)
    with open(os.path.join(OUT, "learning_expected.json"), "w") as f:
        json.dump(graphs, f, indent=0)
    # the raw files too (a few KB): the edgelist writer is compared with them byte for byte (src/io.jl:338-359)
    import shutil
    os.makedirs(os.path.join(OUT, "edgelists"), exist_ok=True)
    for fn in sorted(os.listdir(ld)):
        shutil.copyfile(os.path.join(ld, fn), os.path.join(OUT, "edgelists", fn))
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    sys.exit(main())
