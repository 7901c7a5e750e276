"""The binner (core/binner.cpp, core/src/predictionranges.hh).
CPU: the oracle restatement (oracle/binner_oracle.cpp) reproduces the Bioboxes files the REAL reference binner
wrote for the committed GFF3 goldens (tests/golden/binner_*.tsv, generator make_golden_binner.py).
GPU: trpa_bin_batch == oracle on the golden inputs and on adversarial random tables (uint16 wrap-around of the
reference's medium_unsigned_int, ties in the majority vote, pruning that empties groups); binner-b200 reproduces
the reference's files byte for byte, line order included."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import binner_util as bu
import golden_util as gu
import oracle_lib as ol

ALL = [(c, v) for c in bu.CASES for v in bu.VARIANTS]


def golden_body(case, variant):
    lines = open(bu.golden_path(case, variant)).readlines()
    k = [i for i, l in enumerate(lines) if l.startswith("@@SequenceID")][0]
    return lines[:k + 1], lines[k + 1:]


@pytest.mark.parametrize("case,variant", ALL)
def test_oracle_binner_matches_reference(case, variant):
    data = gu.case_data(case)
    kw = bu.VARIANTS[variant][1]
    recs, sup, gb, names = bu.flat_tables(gu.golden_lines(case), data, kw.get("glob", "(.+)"))
    pp, rank_of_node, pid = bu.params_of(data, **kw)
    res, st = bu.oracle_bin(data, recs, sup, gb, pp, rank_of_node, pid)
    header, body = golden_body(case, variant)
    assert sorted(bu.body_lines(names, res, data)) == sorted(body)
    assert header[4] == "@SampleID:sample_%s\n" % case


def random_tables(rng, data, n_groups, big):
    """Prediction records with arbitrary ranges and per-level supports (not only what taxator prints)."""
    parent, left, right, depth = data.nested_set()
    n = len(parent)
    recs, sup, gb = [], [], [0]
    qid = 0
    for g in range(n_groups):
        k = int(rng.integers(1, 40 if big else 6))
        # most records of a group sit under one species' lineage, some elsewhere
        home = int(rng.integers(1, n))
        for _ in range(k):
            lower = home if rng.random() < 0.6 else int(rng.integers(0, n))
            up_steps = int(rng.integers(0, int(depth[lower]) + 1))
            upper = lower
            for _ in range(up_steps):
                upper = int(parent[upper])
            m = int(depth[lower]) - int(depth[upper]) + 1
            hi = 70000 if big else 900
            s = np.sort(rng.integers(0, hi, m))[::-1] if rng.random() < 0.5 else rng.integers(0, hi, m)
            if rng.random() < 0.3:
                qid += 1
            recs.append((lower, upper, len(sup), int(rng.integers(1, 100000)), qid, 0))
            sup.extend(int(x) for x in s)
        qid += 1
        gb.append(len(recs))
    return np.array(recs, bu.BIN_RECORD), np.array(sup, np.uint32), np.array(gb, np.uint32)


@pytest.mark.parametrize("case,variant", ALL)
def test_host_compiled_kernel_bodies_match_oracle_on_goldens(case, variant):
    data = gu.case_data(case)
    kw = bu.VARIANTS[variant][1]
    recs, sup, gb, names = bu.flat_tables(gu.golden_lines(case), data, kw.get("glob", "(.+)"))
    pp, rank_of_node, pid = bu.params_of(data, **kw)
    want, sw = bu.oracle_bin(data, recs, sup, gb, pp, rank_of_node, pid)
    got, sg = bu.oracle_bin(data, recs, sup, gb, pp, rank_of_node, pid, lib=bu.host_binner())
    same(got, want, sg, sw)


@pytest.mark.parametrize("seed,big", [(1, False), (2, True), (3, True), (4, False)])
def test_host_compiled_kernel_bodies_match_oracle_on_random_tables(seed, big):
    data = gu.case_data("nt_1kb")
    rng = np.random.default_rng(seed)
    recs, sup, gb = random_tables(rng, data, 3000, big)
    modes = set()
    for kw in (dict(), dict(majority=0.5, min_support=1), dict(sample_min="0.01"), dict(sample_min="40000" if big else "500"),
               dict(pid={"species": 0.5, "genus": 0.1, "phylum": 0.01}), dict(majority=0.99, min_support=60000)):
        pp, rank_of_node, pid = bu.params_of(data, **kw)
        want, sw = bu.oracle_bin(data, recs, sup, gb, pp, rank_of_node, pid)
        got, sg = bu.oracle_bin(data, recs, sup, gb, pp, rank_of_node, pid, lib=bu.host_binner())
        same(got, want, sg, sw)
        modes |= set(int(m) for m in want["mode"])
    assert modes >= {1, 2, 3}


def gpu_bin(ctx, data, recs, sup, gb, pp, rank_of_node, pid):
    parent, left, right, depth = data.nested_set()
    ctx.load_taxonomy(parent, left, right, depth, 0)
    out = np.zeros(len(gb) - 1, bu.BIN_RESULT)
    st = bu.BinStats()
    vp = ctypes.c_void_p
    rc = ctx.L.trpa_bin_batch(ctx.h, ctypes.byref(pp), vp(recs.ctypes.data), ctypes.c_uint32(len(recs)), vp(sup.ctypes.data),
                              ctypes.c_uint32(len(sup)), vp(gb.ctypes.data), ctypes.c_uint32(len(gb) - 1),
                              vp(rank_of_node.ctypes.data) if rank_of_node is not None else None,
                              vp(pid.ctypes.data) if pid is not None else None, vp(out.ctypes.data), ctypes.byref(st))
    assert rc == 0, ctx.L.trpa_last_error()
    return out, st


def same(a, b, sa, sb):
    for f in bu.BIN_RESULT.names:
        if f in ("node", "support", "length", "mode"):
            assert np.array_equal(a[f], b[f]), f
    live = a["mode"] != 0
    for f in ("lower_node", "upper_node", "lower_support", "upper_support"):
        assert np.array_equal(a[f][live], b[f][live]), f
    for f in ("nested_taxa", "root_support", "pruned_taxa", "min_support_found"):
        assert getattr(sa, f) == getattr(sb, f), f


@pytest.mark.gpu
@pytest.mark.parametrize("case,variant", ALL)
def test_gpu_binner_matches_oracle_on_goldens(ctx, case, variant):
    data = gu.case_data(case)
    kw = bu.VARIANTS[variant][1]
    recs, sup, gb, names = bu.flat_tables(gu.golden_lines(case), data, kw.get("glob", "(.+)"))
    pp, rank_of_node, pid = bu.params_of(data, **kw)
    want, sw = bu.oracle_bin(data, recs, sup, gb, pp, rank_of_node, pid)
    got, sg = gpu_bin(ctx, data, recs, sup, gb, pp, rank_of_node, pid)
    same(got, want, sg, sw)
    assert sorted(bu.body_lines(names, got, data)) == sorted(golden_body(case, variant)[1])


@pytest.mark.gpu
@pytest.mark.parametrize("seed,big", [(1, False), (2, True), (3, True), (4, False)])
def test_gpu_binner_matches_oracle_on_random_tables(ctx, seed, big):
    data = gu.case_data("nt_1kb")
    rng = np.random.default_rng(seed)
    recs, sup, gb = random_tables(rng, data, 3000, big)
    for kw in (dict(), dict(majority=0.5, min_support=1), dict(sample_min="0.01"), dict(sample_min="40000" if big else "500"),
               dict(pid={"species": 0.5, "genus": 0.1, "phylum": 0.01}), dict(majority=0.99, min_support=60000)):
        pp, rank_of_node, pid = bu.params_of(data, **kw)
        want, sw = bu.oracle_bin(data, recs, sup, gb, pp, rank_of_node, pid)
        got, sg = gpu_bin(ctx, data, recs, sup, gb, pp, rank_of_node, pid)
        same(got, want, sg, sw)
    # empty sample
    pp, _, _ = bu.params_of(data)
    got, sg = gpu_bin(ctx, data, recs[:0], sup[:0], np.zeros(1, np.uint32), pp, None, None)
    assert len(got) == 0 and sg.root_support == 0


@pytest.mark.gpu
def test_gpu_binner_rejects_bad_tables(ctx):
    data = gu.case_data("nt_small")
    recs, sup, gb, names = bu.flat_tables(gu.golden_lines("nt_small"), data)
    pp, _, _ = bu.params_of(data)
    parent, left, right, depth = data.nested_set()
    ctx.load_taxonomy(parent, left, right, depth, 0)
    vp = ctypes.c_void_p
    out = np.zeros(len(gb) - 1, bu.BIN_RESULT)

    def call(r, s, g):
        return ctx.L.trpa_bin_batch(ctx.h, ctypes.byref(pp), vp(r.ctypes.data), ctypes.c_uint32(len(r)), vp(s.ctypes.data),
                                    ctypes.c_uint32(len(s)), vp(g.ctypes.data), ctypes.c_uint32(len(g) - 1), None, None,
                                    vp(out.ctypes.data), None)
    bad = recs.copy()
    k = int(np.flatnonzero(bad["lower_node"] != 0)[0])
    bad["upper_node"][k], bad["lower_node"][k] = bad["lower_node"][k], 0      # upper below lower
    assert call(bad, sup, gb) != 0
    bad = recs.copy()
    bad["support_begin"][-1] = len(sup)
    assert call(bad, sup, gb) != 0
    g2 = gb.copy()
    g2[-1] -= 1
    assert call(recs, sup, g2) != 0
    assert call(recs, sup, gb) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("case", bu.CASES)
def test_binner_cli_matches_reference_bytes(case, tmp_path):
    """binner-b200 on the files the reference binner reads: identical output, line order included."""
    exe = os.path.join(ol.ROOT, "taxator-tk_b200", "bin", "binner-b200")
    assert os.path.exists(exe), "build first: make -C taxator-tk_b200"
    data = gu.case_data(case)
    d = str(tmp_path)
    bu.deep_taxonomy_files(data, d)
    gff = "##gff-version 3\n" + "".join(gu.golden_lines(case))
    env = dict(os.environ, TAXATORTK_TAXONOMY_NCBI=d)
    for variant, (args, _) in bu.VARIANTS.items():
        p = subprocess.run([exe, "-n", "sample_" + case, "-l", os.path.join(d, "binning.log")] + args, cwd=d, env=env,
                           input=gff.encode(), stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert p.returncode == 0, p.stderr.decode()
        assert p.stdout == open(bu.golden_path(case, variant), "rb").read(), variant
    # the same sample split over two files + stdin
    half = len(gu.golden_lines(case)) // 2
    open(os.path.join(d, "a.gff3"), "w").write("##gff-version 3\n" + "".join(gu.golden_lines(case)[:half]))
    p = subprocess.run([exe, "-n", "sample_" + case, "-l", os.path.join(d, "binning.log"), "-f", "a.gff3", "-"], cwd=d, env=env,
                       input=("".join(gu.golden_lines(case)[half:])).encode(), stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0, p.stderr.decode()
    assert p.stdout == open(bu.golden_path(case, "default"), "rb").read()


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["nt_small", "nt_1kb"])
def test_taxator_b200_piped_into_binner_b200_matches_reference_pipeline(case, tmp_path):
    """`taxator-b200 | binner-b200` against `taxator -p 1 | binner` of the reference (tests/golden/binner_<case>_pipeline_*.tsv):
    the same bytes, line order included -- both taxators print their records in input order."""
    taxator = os.path.join(ol.ROOT, "taxator-tk_b200", "bin", "taxator-b200")
    binner = os.path.join(ol.ROOT, "taxator-tk_b200", "bin", "binner-b200")
    data = gu.case_data(case)
    d = str(tmp_path)
    bu.deep_taxonomy_files(data, d)
    env = dict(os.environ, TAXATORTK_TAXONOMY_NCBI=d)
    with open(os.path.join(d, "alignments.tsv"), "rb") as fin:
        p = subprocess.run([taxator, "-a", "rpa", "-g", "mapping.tax", "-q", "query.fna", "-f", "ref.fna", "-i", "ref.fna.fai"], cwd=d, env=env,
                           stdin=fin, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode == 0, p.stderr.decode()
    gff = p.stdout
    for variant in ("default", "glob10"):
        p = subprocess.run([binner, "-n", "sample_" + case, "-l", os.path.join(d, "binning.log")] + bu.VARIANTS[variant][0], cwd=d, env=env,
                           input=gff, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert p.returncode == 0, p.stderr.decode()
        assert p.stdout == open(bu.golden_path(case, "pipeline_" + variant), "rb").read(), variant
