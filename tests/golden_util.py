"""Helpers to compare prediction results with the committed golden GFF3 of the real reference."""
import json
import os

import oracle_lib as ol
import synth
import gff3

GOLDEN = os.path.join(ol.ROOT, "tests", "golden")
CASES = json.load(open(os.path.join(GOLDEN, "cases.json")))
try:
    HEAVY = tuple(json.load(open(os.path.join(GOLDEN, "cases_meta.json")))["heavy"])
except Exception:
    HEAVY = ()


def long_pair(seed, L, rate, indel, n_over_m=1.0):
    """Deterministic long nucleotide pair (only the seed is stored in the fixtures): a = uniform ACGT of length
    L; b = an independent window-free mutation of a (rate substitutions; if indel, a third of the events are
    deletions and a third insertions), optionally extended by random bases to n_over_m * L."""
    import numpy as np
    rng = np.random.default_rng(seed)
    nt = np.frombuffer(b"ACGT", np.uint8)
    a = nt[rng.integers(0, 4, L)]
    r = rng.random(L)
    b = a.copy()
    sub = r < (rate / 3 if indel else rate)
    b[sub] = nt[(np.searchsorted(nt, b[sub]) + rng.integers(1, 4, int(sub.sum()))) % 4]
    if indel:
        reps = np.ones(L, np.int64)
        reps[(r >= rate / 3) & (r < 2 * rate / 3)] = 0
        ins = (r >= 2 * rate / 3) & (r < rate)
        reps[ins] = 2
        b = np.repeat(b, reps)
        pos = np.cumsum(reps)[ins] - 1
        b[pos] = nt[rng.integers(0, 4, len(pos))]
    extra = int(round(L * (n_over_m - 1.0)))
    if extra > 0:
        b = np.concatenate([b, nt[rng.integers(0, 4, extra)]])
    return np.ascontiguousarray(a), np.ascontiguousarray(b)



def long_pairs():
    """[(a, b, distance computed by the real SeqAn)] of tests/golden/seqan_long_pairs.json."""
    out = []
    for spec in json.load(open(os.path.join(GOLDEN, "seqan_long_pairs.json"))):
        a, b = long_pair(spec["seed"], spec["L"], spec["rate"], spec["indel"], spec["n_over_m"])
        assert len(a) == spec["m"] and len(b) == spec["n"]
        out.append((a, b, spec["dist"]))
    return out


def case_data(name):
    d = dict(CASES[name])
    d["query_len"] = tuple(d["query_len"])
    d["levels"] = tuple(d["levels"])
    return synth.generate(synth.SynthConfig(**d))


def golden_lines(name):
    return open(os.path.join(GOLDEN, name + ".gff3")).readlines()


def render_sorted(fd, results):
    d = fd.d
    taxids = [str(t) for t in d.tax_ids]
    lines = gff3.render(results, fd.segs, d.q_names, fd.q_len, fd.parent, fd.depth, taxids)
    return sorted(lines)


# ---------------------------------------------------------------- alignment-free models (simple-lca, megan-lca, ...)
# Variants of the golden cases for the models of core/src/taxonpredictionmodel.hh:57-259: scores quantised so that
# ties occur (best-score sets, n-best distinct values), an e-value column that varies, and some taxa named
# "unclassified ..." (inherited by their descendants, ncbidata.cpp:119-126).  Used by make_golden_lca.py (reference
# binary -> tests/golden/lca_<case>_<variant>.gff3) and by the tests, so both see the same inputs.
LCA_CASES = ("nt_small", "nt_1kb")
LCA_VARIANTS = {
    "dummy": (["-a", "dummy"], dict(model=0)),
    "simple": (["-a", "simple-lca"], dict(model=1)),
    "megan_default": (["-a", "megan-lca"], dict(model=2, toppercent=0.05, minscore=0.0, maxevalue=1000.0, minsupport=1)),
    # -c stays 1 in the golden runs: the filter's support counts the rises of the running maximum in record-set
    # order, and the reference orders records of equal (qstart, qstop) by HEAP ADDRESS (std::sort on a tuple whose
    # third member is the record pointer, alignmentrecord.hh:479), so its -c >= 2 output is allocator dependent
    "megan_strict": (["-a", "megan-lca", "-t", "0.3", "-m", "120", "-e", "1e-9", "-c", "1"],
                     dict(model=2, toppercent=0.3, minscore=120.0, maxevalue=1e-9, minsupport=1)),
    "megan_uncl": (["-a", "ic-megan-lca", "-t", "0.5", "-u"],
                   dict(model=2, toppercent=0.5, minscore=0.0, maxevalue=1000.0, minsupport=1, ignore_unclassified=True)),
    "nbest1": (["-a", "n-best-lca", "-n", "1"], dict(model=3, nbest=1)),
    "nbest3": (["-a", "n-best-lca", "-n", "3"], dict(model=3, nbest=3)),
}


# megan-lca with -c >= 2: the support filter counts rises of the running maximum in record-set ORDER, and the reference
# orders records of equal (qstart, qstop) by heap address (alignmentrecord.hh:479).  On inputs whose records have
# DISTINCT query ranges inside a query the order is (qstart, qstop) for both sides, so -c >= 2 can be pinned there.
LCA_DISTINCT_CASE = "nt_small"
LCA_C_VARIANTS = {
    "megan_c2": (["-a", "megan-lca", "-c", "2"], dict(model=2, toppercent=0.05, minscore=0.0, maxevalue=1000.0, minsupport=2)),
    "megan_c3_t50": (["-a", "megan-lca", "-c", "3", "-t", "0.5"], dict(model=2, toppercent=0.5, minscore=0.0, maxevalue=1000.0, minsupport=3)),
    "megan_c4_t100_u": (["-a", "ic-megan-lca", "-c", "4", "-t", "1.0", "-u"],
                        dict(model=2, toppercent=1.0, minscore=0.0, maxevalue=1000.0, minsupport=4, ignore_unclassified=True)),
}


def lca_distinct_case_data():
    """lca_case_data(LCA_DISTINCT_CASE) with the query ranges of a query's records made pairwise distinct (qstop
    shortened by the record's rank among those that share its range)."""
    import numpy as np
    d, evalue, named = lca_case_data(LCA_DISTINCT_CASE)
    r = d.rec
    seen = {}
    for k in range(len(r["q"])):
        key = (int(r["q"][k]), int(r["qstart"][k]), int(r["qstop"][k]))
        while key in seen and key[2] > key[1]:
            key = (key[0], key[1], key[2] - 1)
        assert key not in seen
        seen[key] = k
        r["qstop"][k] = key[2]
    return d, evalue, named


def lca_masked(d):
    """records masked in the input ('*' prefix, AlignmentRecord::isFiltered): every 17th line and ALL lines of
    the fourth query (a record set without any active record)"""
    import numpy as np
    n = len(d.rec["q"])
    return (np.arange(n) % 17 == 5) | (d.rec["q"] == 3)


def lca_case_data(name):
    """(SynthData with quantised scores, evalue per record in FILE order, unclassified-name flag per node)."""
    import numpy as np
    d = case_data(name)
    sc = d.rec["score"].astype(np.float64)
    d.rec["score"] = (np.round(sc / 8.0) * 8.0).astype(np.float32)
    evalue = np.power(10.0, -np.round(d.rec["score"].astype(np.float64) / 12.0))
    named = np.array([(i % 11) == 5 and i != 0 for i in range(len(d.tax_ids))], bool)
    return d, evalue, named


def lca_unclassified_flags(d, named):
    """Taxon::is_unclassified: own name or any ancestor's name (the root excepted) contains 'unclassified'."""
    import numpy as np
    n = len(d.tax_ids)
    out = np.zeros(n, np.uint8)
    for i in range(1, n):          # parents precede children in the synthetic taxonomy (index order)
        p = int(d.tax_parent[i])
        assert p < i
        out[i] = 1 if (named[i] or out[p]) else 0
    return out


def lca_write_files(d, evalue, named, outdir):
    """nodes/names/mapping/alignments as the reference's taxator reads them (no sequence files needed)."""
    os.makedirs(outdir, exist_ok=True)
    d.write_files(outdir)
    with open(os.path.join(outdir, "names.dmp"), "w") as f:
        for i, t in enumerate(d.tax_ids):
            f.write("%d\t|\t%snode%d\t|\t\t|\tscientific name\t|\n" % (t, "unclassified " if named[i] else "", t))
    r = d.rec
    masked = lca_masked(d)
    with open(os.path.join(outdir, "alignments.tsv"), "w") as f:
        for k in range(len(r["q"])):
            qi = r["q"][k]
            f.write("%s%s\t%d\t%d\t%d\t%s\t%d\t%d\t%s\t%s\t%d\t%d\n" % (
                "*" if masked[k] else "",
                d.q_names[qi], r["qstart"][k], r["qstop"][k], len(d.q_seqs[qi]), d.ref_names[r["r"][k]], r["rstart"][k],
                r["rstop"][k], repr(float(r["score"][k])), repr(float(evalue[k])), r["ident"][k], r["alnlen"][k]))


def lca_flat(d, evalue):
    """segment / candidate tables + evalue in CANDIDATE order (same ordering as SynthData.segments())."""
    import numpy as np
    r = d.rec
    nrec = len(r["q"])
    order = np.lexsort((np.arange(nrec), r["qstop"], r["qstart"], r["q"]))
    segs, cands = d.segments()           # record sets are formed from ALL records, masked or not
    keep = ~lca_masked(d)[order]         # ... and only the unmasked ones go into the candidate table
    ev = evalue[order]
    csum = np.concatenate([[0], np.cumsum(keep)])
    out = segs.copy()
    for i, sg in enumerate(segs):
        b, c = int(sg["cand_begin"]), int(sg["cand_count"])
        out[i]["cand_begin"] = csum[b]
        out[i]["cand_count"] = csum[b + c] - csum[b]
    return out, np.ascontiguousarray(cands[keep]), np.ascontiguousarray(ev[keep])


def lca_golden_lines(case, variant):
    return open(os.path.join(GOLDEN, "lca_%s_%s.gff3" % (case, variant))).readlines()
