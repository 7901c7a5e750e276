"""Edge-case inputs shared by the CPU (host state machine) and GPU parity tests."""
import numpy as np

import oracle_lib as ol
import synth


def base(seed=11, protein=False, **kw):
    cfg = dict(seed=seed, protein=protein, n_genomes=40, genome_len=3000 if not protein else 500, n_queries=80,
               query_len=(150, 700) if not protein else (60, 250), n_cand=18, levels=(2, 3, 5, 8, 12))
    cfg.update(kw)
    return ol.FlatData(synth.generate(synth.SynthConfig(**cfg)))


def ranges_past_ends(protein, far=True):
    """Reference coordinates beyond the stored sequence are clipped (sequencestorage.hh:353,
    faidx.h:325-331); a start past the end yields an empty segment (distance = other length).
    far=False leaves the start-past-the-end records out: on an EMPTY segment the real reference segfaults in the
    nucleotide path and, in the protein path, gets INT_MIN from SeqAn for every alignment with an empty string, overflows
    2 * INT_MIN and ends up with distance 0 (undefined behaviour; here an empty segment is at distance = the other length)."""
    fd = base(seed=21, protein=protein)
    rng = np.random.default_rng(5)
    c = fd.cands
    L = fd.r_len[c["ref_seq"]].astype(np.int64)
    pick = rng.random(len(c)) < 0.25
    fwd = c["rstart"] <= c["rstop"]
    c["rstop"][pick & fwd] = (L[pick & fwd] + rng.integers(1, 500, int((pick & fwd).sum()))).astype(np.uint32)
    c["rstart"][pick & ~fwd] = (L[pick & ~fwd] + rng.integers(1, 500, int((pick & ~fwd).sum()))).astype(np.uint32)
    far = (rng.random(len(c)) < 0.03) & bool(far)
    c["rstart"][far & fwd] = (L[far & fwd] + 10).astype(np.uint32)
    c["rstop"][far & fwd] = (L[far & fwd] + 200).astype(np.uint32)
    return fd


def n_rich(letters=b"NRYKMnacgt-"):
    """letters: what replaces 5 % of the query bases (the real reference's FASTA reader only accepts ACGTN in either case)."""
    fd = base(seed=22, frac_n=0.08)
    q = fd.q_chars.copy()
    rng = np.random.default_rng(1)
    m = rng.random(len(q)) < 0.05
    q[m] = np.frombuffer(letters, np.uint8)[rng.integers(0, len(letters), int(m.sum()))]
    fd.q_chars = q
    fd.q_codes = ol.codes_of(q, False)
    return fd


def special_segments():
    fd = base(seed=23)
    segs, c = fd.segs, fd.cands
    segs["cand_count"][0] = 1                      # n == 1 (hh:371-388)
    segs["cand_count"][1] = 0                      # n == 0 (hh:359-368)
    s = segs[2]                                    # 100% full-length top hit + a score tie (hh:431-472)
    b, n = int(s["cand_begin"]), int(s["cand_count"])
    qs, qe = c["qstart"][b:b + n].min(), c["qstop"][b:b + n].max()
    L = int(qe - qs + 1)
    c["qstart"][b], c["qstop"][b] = qs, qe
    c["alnlen"][b] = L; c["identities"][b] = L; c["score"][b] = 1e6
    c["score"][b + 1] = 1e6
    s = segs[3]                                    # identical hit that is NOT the best score (hh:511-517)
    b, n = int(s["cand_begin"]), int(s["cand_count"])
    qs, qe = c["qstart"][b:b + n].min(), c["qstop"][b:b + n].max()
    L = int(qe - qs + 1)
    worst = b + int(np.argmin(c["score"][b:b + n]))
    c["qstart"][worst], c["qstop"][worst] = qs, qe
    c["alnlen"][worst] = L; c["identities"][worst] = L
    return fd


def many_candidates():
    cfg = synth.SynthConfig(seed=25, n_genomes=300, genome_len=1500, n_queries=6, query_len=(400, 600), n_cand=300,
                            levels=(3, 6, 12, 30, 80))
    return ol.FlatData(synth.generate(cfg))


def score_ties():
    """Many equal (score, identities) keys: SortFilter is a STABLE sort (alignmentsfilter.hh:171-190),
    so record order decides -- the device-side rank sort must reproduce it."""
    fd = base(seed=26, n_queries=120)
    c = fd.cands
    c["score"] = (np.floor(c["score"] / 60.0) * 60.0).astype(np.float32)
    c["identities"] = (c["identities"] // 16) * 16
    c["identities"] = np.minimum(c["identities"], c["alnlen"] - 1).astype(np.uint32)
    return fd


def write_files_from_flat(fd, outdir):
    """The files the reference reads, from the FLAT tables (the edge cases edit those, not the SynthData records):
    taxonomy / mapping / reference FASTA of fd.d, query FASTA from fd.q_chars, one alignment line per candidate in table
    order (inside a query that is (qstart, qstop, file order), which the reference's record-set generator keeps)."""
    import os
    d = fd.d
    d.write_files(outdir)
    q_seqs = [fd.q_chars[int(o):int(o) + int(l)] for o, l in zip(fd.q_off, fd.q_len)]
    d._write_fasta(os.path.join(outdir, "query.fna"), d.q_names, q_seqs, fai=False)
    with open(os.path.join(outdir, "alignments.tsv"), "w") as f:
        for sg in fd.segs:
            q = int(sg["query_seq"])
            for c in fd.cands[int(sg["cand_begin"]):int(sg["cand_begin"]) + int(sg["cand_count"])]:
                f.write("%s\t%d\t%d\t%d\t%s\t%d\t%d\t%s\t0\t%d\t%d\n" % (
                    d.q_names[q], c["qstart"], c["qstop"], fd.q_len[q], d.ref_names[int(c["ref_seq"])], c["rstart"], c["rstop"],
                    repr(float(c["score"])), c["identities"], c["alnlen"]))


PARAM_SETS = [("0.0", "0.05"), ("0.9", "0.05"), ("0.5", "0.3"), ("0.5", "0.0")]   # (-x, -t) as the command line spells them


def reference_cases():
    """Edge cases that are also run through the REAL reference binary (tests/golden/make_golden_edge.py)."""
    return [("past_ends_nt", ranges_past_ends(False, far=False)), ("past_ends_aa", ranges_past_ends(True, far=False)),
            ("n_rich", n_rich(b"NNnnacgt")), ("many_candidates", many_candidates())]
