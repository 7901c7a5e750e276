"""Synthetic refpack / taxonomy / query / alignment generator for the RPA hot path.

Follows SURVEY.md section 8(d): a random taxonomy with named ranks from the reference's
``default_ranks`` (core/src/constants.hh:32), genomes evolved down that tree by substitutions,
queries cut from genomes with extra noise, and synthetic alignment records (the 12-column TAB
format of doc/fileformats.md:9-26) for the top-K most similar genomes over the same window.
Everything is derived from one integer seed.  The data can be written as the files the reference
binary consumes (nodes.dmp, names.dmp, FASTA + .fai, mapping, alignments) and flattened into the
arrays the C ABI consumes.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

RANKS = ["superkingdom", "phylum", "class", "order", "family", "genus", "species"]
NT = np.frombuffer(b"ACGT", dtype=np.uint8)
AA20 = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[:] = ord("N")
for _a, _b in zip(b"ACGTacgt", b"TGCAtgca"):
    _COMP[_a] = _b


def revcomp(a: np.ndarray) -> np.ndarray:
    return _COMP[a[::-1]]


@dataclass
class SynthConfig:
    seed: int = 1
    protein: bool = False
    n_genomes: int = 200
    genome_len: int = 20000
    n_queries: int = 1000
    query_len: tuple = (1000, 1000)  # inclusive range
    n_cand: int = 36                 # top-K genomes per query
    levels: tuple = (4, 8, 16, 40, 100)  # node counts for the ranks above species (6 ranks with species)
    edge_rate: tuple = (0.01, 0.04)  # per-edge substitution rate range
    query_sub: float = 0.05          # extra substitutions in the query
    query_indel: float = 0.0         # extra insertions+deletions in the query (long-read mode)
    frac_rev_genomes: float = 0.3    # genomes stored reverse-complemented (nucleotide only)
    frac_partial: float = 0.25       # records covering only part of the query (exercise left/right ext)
    frac_n: float = 0.0              # fraction of reference bases replaced by N
    frac_trunc_genomes: float = 0.1  # genomes stored truncated (exercise clipping at both ends)
    line_width: int = 70
    multi_segment_frac: float = 0.0  # queries whose records fall into two disjoint segments


@dataclass
class SynthData:
    cfg: SynthConfig
    # taxonomy (before the reference's pruning: every node has one of the 7 ranks or is the root)
    tax_ids: np.ndarray = None       # int taxids, tax_ids[0] == 1 (root)
    tax_parent: np.ndarray = None    # index of parent node (root -> itself)
    tax_rank: list = None            # rank name per node ("no rank" for the root)
    # reference store
    ref_names: list = None
    ref_seqs: list = None            # list of uint8 ASCII arrays (as stored)
    ref_taxnode: np.ndarray = None   # node index (species) per reference sequence
    # query store
    q_names: list = None
    q_seqs: list = None
    # alignment records, grouped by query, in file order
    rec: dict = field(default_factory=dict)

    # ------------------------------------------------------------------ flat views for the C ABI
    def nested_set(self):
        """parent/left/right/depth arrays (node index space = tax_ids order), root index 0."""
        n = len(self.tax_ids)
        children = [[] for _ in range(n)]
        for i in range(1, n):
            children[self.tax_parent[i]].append(i)
        left = np.zeros(n, np.uint32)
        right = np.zeros(n, np.uint32)
        depth = np.zeros(n, np.uint8)
        counter = 0
        stack = [(0, 0)]
        left[0] = counter = 1
        it = [0] * n
        path = [0]
        while path:
            v = path[-1]
            if it[v] < len(children[v]):
                c = children[v][it[v]]
                it[v] += 1
                counter += 1
                left[c] = counter
                depth[c] = depth[v] + 1
                path.append(c)
            else:
                counter += 1
                right[v] = counter
                path.pop()
        return self.tax_parent.astype(np.uint32), left, right, depth

    def store_arrays(self, seqs):
        lens = np.array([len(s) for s in seqs], dtype=np.uint32)
        off = np.zeros(len(seqs), dtype=np.uint64)
        if len(seqs) > 1:
            off[1:] = np.cumsum(lens[:-1].astype(np.uint64))
        chars = np.concatenate(seqs) if len(seqs) else np.zeros(0, np.uint8)
        return np.ascontiguousarray(chars), off, lens

    def segments(self):
        """Record sets exactly as RecordSetGeneratorUnsorted<split=true> forms them
        (core/src/alignmentrecord.hh:457-504): per query, records ordered by (qstart, qstop, file
        order) and cut where a record starts after the running maximum stop.
        Returns (segs[n,3]: query index, cand_begin, cand_count; cand structured array)."""
        r = self.rec
        nrec = len(r["q"])
        order = np.lexsort((np.arange(nrec), r["qstop"], r["qstart"], r["q"]))
        q = r["q"][order]
        qs = r["qstart"][order]
        qe = r["qstop"][order]
        segs = []
        begin = 0
        run_stop = 0
        for k in range(nrec):
            if k == 0 or q[k] != q[k - 1]:
                if k:
                    segs.append((q[k - 1], begin, k - begin))
                begin = k
                run_stop = qe[k]
            elif qs[k] > run_stop:
                segs.append((q[k - 1], begin, k - begin))
                begin = k
                run_stop = qe[k]
            else:
                run_stop = max(run_stop, qe[k])
        if nrec:
            segs.append((q[nrec - 1], begin, nrec - begin))
        cand = np.zeros(nrec, dtype=CAND_DTYPE)
        cand["ref_seq"] = r["r"][order]
        cand["rstart"] = r["rstart"][order]
        cand["rstop"] = r["rstop"][order]
        cand["qstart"] = qs
        cand["qstop"] = qe
        cand["score"] = r["score"][order]
        cand["identities"] = r["ident"][order]
        cand["alnlen"] = r["alnlen"][order]
        cand["node"] = self.ref_taxnode[r["r"][order]]
        seg = np.zeros(len(segs), dtype=SEG_DTYPE)
        for i, (qq, b, c) in enumerate(segs):
            seg[i] = (qq, b, c, 0)
        return seg, cand

    # ---------------------------------------------------------------------- files for the reference
    def write_files(self, outdir: str):
        os.makedirs(outdir, exist_ok=True)
        with open(os.path.join(outdir, "nodes.dmp"), "w") as f:
            for i, t in enumerate(self.tax_ids):
                p = self.tax_ids[self.tax_parent[i]]
                f.write("%d\t|\t%d\t|\t%s\t|\t\t|\n" % (t, p, self.tax_rank[i]))
        with open(os.path.join(outdir, "names.dmp"), "w") as f:
            for i, t in enumerate(self.tax_ids):
                f.write("%d\t|\tnode%d\t|\t\t|\tscientific name\t|\n" % (t, t))
        self._write_fasta(os.path.join(outdir, "ref.fna"), self.ref_names, self.ref_seqs, fai=True)
        self._write_fasta(os.path.join(outdir, "query.fna"), self.q_names, self.q_seqs, fai=False)
        with open(os.path.join(outdir, "mapping.tax"), "w") as f:
            for name, node in zip(self.ref_names, self.ref_taxnode):
                f.write("%s\t%d\n" % (name, self.tax_ids[node]))
        r = self.rec
        with open(os.path.join(outdir, "alignments.tsv"), "w") as f:
            for k in range(len(r["q"])):
                qi = r["q"][k]
                f.write("%s\t%d\t%d\t%d\t%s\t%d\t%d\t%s\t0\t%d\t%d\n" % (
                    self.q_names[qi], r["qstart"][k], r["qstop"][k], len(self.q_seqs[qi]),
                    self.ref_names[r["r"][k]], r["rstart"][k], r["rstop"][k],
                    repr(float(r["score"][k])), r["ident"][k], r["alnlen"][k]))

    def _write_fasta(self, path, names, seqs, fai):
        w = self.cfg.line_width
        offset = 0
        fai_lines = []
        with open(path, "wb") as f:
            for name, s in zip(names, seqs):
                hdr = (">%s\n" % name).encode()
                f.write(hdr)
                offset += len(hdr)
                fai_lines.append("%s\t%d\t%d\t%d\t%d\n" % (name, len(s), offset, w, w + 1))
                b = s.tobytes()
                for i in range(0, len(b), w):
                    f.write(b[i:i + w])
                    f.write(b"\n")
                offset += len(b) + (len(b) + w - 1) // w
        if fai:
            with open(path + ".fai", "w") as f:
                f.writelines(fai_lines)


CAND_DTYPE = np.dtype([("ref_seq", "<u4"), ("rstart", "<u4"), ("rstop", "<u4"), ("qstart", "<u4"), ("qstop", "<u4"),
                       ("score", "<f4"), ("identities", "<u4"), ("alnlen", "<u4"), ("node", "<u4")])
SEG_DTYPE = np.dtype([("query_seq", "<u4"), ("cand_begin", "<u4"), ("cand_count", "<u4"), ("reserved", "<u4")])
RESULT_DTYPE = np.dtype([("qrstart", "<u4"), ("qrstop", "<u4"), ("lower_node", "<u4"), ("upper_node", "<u4"),
                         ("rtax_node", "<u4"), ("support", "<u4"), ("ival", "<f4"), ("signal", "<f4"),
                         ("n_pass0", "<u4"), ("n_pass1", "<u4"), ("n_pass2", "<u4"), ("kind", "<u4"),
                         ("cells", "<u8")])


def _mutate_subst(rng, seq, rate, alphabet):
    out = seq.copy()
    m = rng.random(len(seq)) < rate
    k = int(m.sum())
    if k:
        # replace by a *different* letter
        idx = np.searchsorted(alphabet, out[m]) if len(alphabet) > 4 else None
        if idx is None:
            lut = np.zeros(256, np.int64)
            for i, ch in enumerate(alphabet):
                lut[ch] = i
            idx = lut[out[m]]
        out[m] = alphabet[(idx + rng.integers(1, len(alphabet), k)) % len(alphabet)]
    return out


def _mutate_indel(rng, seq, rate, alphabet):
    """rate/2 deletions, rate/2 single-letter insertions (after the position)."""
    if rate <= 0:
        return seq
    r = rng.random(len(seq))
    reps = np.ones(len(seq), np.int64)
    reps[r < rate / 2] = 0
    reps[(r >= rate / 2) & (r < rate)] = 2
    out = np.repeat(seq, reps)
    ends = np.cumsum(reps)
    pos = ends[reps == 2] - 1
    out[pos] = alphabet[rng.integers(0, len(alphabet), len(pos))]
    return out


def generate(cfg: SynthConfig, identity_fn=None) -> SynthData:
    rng = np.random.default_rng(cfg.seed)
    alphabet = AA20 if cfg.protein else NT
    d = SynthData(cfg=cfg)

    # ---- taxonomy: root + len(levels) inner ranks + species
    counts = list(cfg.levels) + [cfg.n_genomes]
    ranks = RANKS[len(RANKS) - len(counts):]
    ids, parent, rank = [1], [0], ["no rank"]
    prev_level = [0]
    next_id = 2
    level_nodes = []
    for lvl, cnt in enumerate(counts):
        cur = []
        for k in range(cnt):
            # first nodes cover every parent once, the rest attach at random
            p = prev_level[k] if k < len(prev_level) else prev_level[int(rng.integers(0, len(prev_level)))]
            ids.append(next_id); parent.append(p); rank.append(ranks[lvl])
            cur.append(len(ids) - 1)
            next_id += 1
        level_nodes.append(cur)
        prev_level = cur
    d.tax_ids = np.array(ids, dtype=np.int64)
    d.tax_parent = np.array(parent, dtype=np.int64)
    d.tax_rank = rank

    # ---- sequences evolved down the tree (substitutions only: coordinates stay aligned)
    G = cfg.genome_len
    node_seq = {0: alphabet[rng.integers(0, len(alphabet), G)]}
    for lvl, nodes in enumerate(level_nodes):
        for v in nodes:
            rate = rng.uniform(*cfg.edge_rate)
            node_seq[v] = _mutate_subst(rng, node_seq[int(d.tax_parent[v])], rate, alphabet)
        if lvl > 0:
            for v in level_nodes[lvl - 1]:
                node_seq.pop(v, None)
    species = level_nodes[-1]
    evolved = [node_seq[v] for v in species]   # forward, full-length, coordinate-aligned
    # order genomes by tree position so that neighbours in index space are relatives
    key = []
    for v in species:
        path = []
        x = v
        while x != 0:
            path.append(x)
            x = int(d.tax_parent[x])
        key.append(tuple(reversed(path)))
    order = sorted(range(len(species)), key=lambda i: key[i])
    evolved = [evolved[i] for i in order]
    species = [species[i] for i in order]

    n_g = cfg.n_genomes
    is_rev = (rng.random(n_g) < cfg.frac_rev_genomes) & (not cfg.protein)
    trunc = rng.random(n_g) < cfg.frac_trunc_genomes
    t_a = np.where(trunc, rng.integers(0, max(1, G // 10), n_g), 0)
    t_b = np.where(trunc, G - rng.integers(0, max(1, G // 10), n_g), G)
    d.ref_names, d.ref_seqs = [], []
    for g in range(n_g):
        s = evolved[g][t_a[g]:t_b[g]]
        if cfg.frac_n > 0:
            s = s.copy()
            s[rng.random(len(s)) < cfg.frac_n] = ord("N") if not cfg.protein else ord("X")
        if is_rev[g]:
            s = revcomp(s)
        d.ref_names.append("G%05d" % g)
        d.ref_seqs.append(np.ascontiguousarray(s))
    d.ref_taxnode = np.array(species, dtype=np.uint32)

    # ---- queries and records
    K = min(cfg.n_cand, n_g)
    d.q_names, d.q_seqs = [], []
    R = {k: [] for k in ("q", "qstart", "qstop", "r", "rstart", "rstop", "score", "ident", "alnlen")}
    for qi in range(cfg.n_queries):
        L = int(rng.integers(cfg.query_len[0], cfg.query_len[1] + 1))
        L = min(L, G)
        g0 = int(rng.integers(0, n_g))
        p = int(rng.integers(0, G - L + 1))
        src = evolved[g0][p:p + L]
        qseq = _mutate_subst(rng, src, cfg.query_sub, alphabet)
        qseq = _mutate_indel(rng, qseq, cfg.query_indel, alphabet)
        d.q_names.append("Q%06d" % qi)
        d.q_seqs.append(np.ascontiguousarray(qseq))
        Lq = len(qseq)
        # candidate genomes: the K nearest in tree order around g0
        lo = max(0, min(g0 - K // 2, n_g - K))
        cands = np.arange(lo, lo + K)
        # ungapped identity of the *source window* (coordinates are aligned) -- good enough as the
        # aligner-reported identity; the query's own noise is applied as an expected-value scale.
        win = np.stack([evolved[g][p:p + L] for g in cands])
        ident_full = (win == src[None, :])
        split2 = rng.random() < cfg.multi_segment_frac and L >= 200
        recs = []
        for ci, g in enumerate(cands):
            qs, qe = 1, L
            if split2:
                # two disjoint groups of records: first 40% / last 40% of the query
                if ci % 2 == 0:
                    qs, qe = 1, int(L * 0.4)
                else:
                    qs, qe = int(L * 0.6), L
            if rng.random() < cfg.frac_partial:
                w = qe - qs + 1
                a = int(rng.integers(0, max(1, w // 4)))
                b = int(rng.integers(0, max(1, w // 4)))
                qs, qe = qs + a, qe - b
            # clip to what the (possibly truncated) stored genome covers
            fs, fe = p + qs, p + qe                   # 1-based forward positions on the evolved genome
            fs_c, fe_c = max(fs, t_a[g] + 1), min(fe, t_b[g])
            if fe_c - fs_c + 1 < 20:
                continue
            qs += fs_c - fs
            qe -= fe - fe_c
            fs, fe = fs_c, fe_c
            ident = int(ident_full[ci, qs - 1:qe].sum())
            ident = int(round(ident * (1.0 - cfg.query_sub)))
            alnlen = qe - qs + 1
            # stored coordinates
            s1, e1 = fs - t_a[g], fe - t_a[g]
            if is_rev[g]:
                glen = t_b[g] - t_a[g]
                rstart, rstop = glen - s1 + 1, glen - e1 + 1   # rstart > rstop: reverse strand
            else:
                rstart, rstop = s1, e1
            # with indels the query coordinates are rescaled to the mutated query
            if Lq != L:
                qs = max(1, min(Lq, int(round(qs * Lq / L))))
                qe = max(qs, min(Lq, int(round(qe * Lq / L))))
            recs.append([qi, qs, qe, int(g), int(rstart), int(rstop), ident, alnlen])
        if not recs:
            continue
        # tie-free (score, identities) inside a query: strictly decreasing identities
        recs.sort(key=lambda r: -r[6])
        for k in range(1, len(recs)):
            recs[k][6] = min(recs[k][6], recs[k - 1][6] - 1)
        recs = [r for r in recs if r[6] > 0]
        perm = rng.permutation(len(recs))   # file order is not score order
        for k in perm:
            qi_, qs, qe, g, rstart, rstop, ident, alnlen = recs[k]
            ident = min(ident, alnlen - 1)  # never a 100% full-length hit (hh:431 shortcut)
            R["q"].append(qi_); R["qstart"].append(qs); R["qstop"].append(qe); R["r"].append(g)
            R["rstart"].append(rstart); R["rstop"].append(rstop); R["ident"].append(ident); R["alnlen"].append(alnlen)
            R["score"].append(float(2 * ident - 3 * (alnlen - ident)))
    d.rec = {
        "q": np.array(R["q"], np.uint32), "qstart": np.array(R["qstart"], np.uint32),
        "qstop": np.array(R["qstop"], np.uint32), "r": np.array(R["r"], np.uint32),
        "rstart": np.array(R["rstart"], np.uint32), "rstop": np.array(R["rstop"], np.uint32),
        "score": np.array(R["score"], np.float32), "ident": np.array(R["ident"], np.uint32),
        "alnlen": np.array(R["alnlen"], np.uint32),
    }
    return d


class FastReference:
    """Taxonomy + refpack + block-resolution identity tables of generate_fast, built once from cfg.seed."""
    pass


def build_reference_fast(cfg: SynthConfig, rng, block: int = 100):
    """First half of generate_fast: taxonomy, evolved genomes, the stored refpack and the cumulative block
    identities between neighbouring genomes.  Returns (SynthData without queries, FastReference)."""
    alphabet = AA20 if cfg.protein else NT
    d = SynthData(cfg=cfg)
    counts = list(cfg.levels) + [cfg.n_genomes]
    ranks = RANKS[len(RANKS) - len(counts):]
    ids, parent, rank = [1], [0], ["no rank"]
    prev_level = [0]
    next_id = 2
    level_nodes = []
    for lvl, cnt in enumerate(counts):
        cur = []
        for k in range(cnt):
            p = prev_level[k] if k < len(prev_level) else prev_level[int(rng.integers(0, len(prev_level)))]
            ids.append(next_id); parent.append(p); rank.append(ranks[lvl])
            cur.append(len(ids) - 1)
            next_id += 1
        level_nodes.append(cur)
        prev_level = cur
    d.tax_ids = np.array(ids, dtype=np.int64)
    d.tax_parent = np.array(parent, dtype=np.int64)
    d.tax_rank = rank
    G = (cfg.genome_len // block) * block
    node_seq = {0: alphabet[rng.integers(0, len(alphabet), G)]}
    for lvl, nodes in enumerate(level_nodes):
        for v in nodes:
            node_seq[v] = _mutate_subst(rng, node_seq[int(d.tax_parent[v])], rng.uniform(*cfg.edge_rate), alphabet)
        if lvl > 0:
            for v in level_nodes[lvl - 1]:
                node_seq.pop(v, None)
    species = level_nodes[-1]
    key = []
    for v in species:
        path = []
        x = v
        while x != 0:
            path.append(x)
            x = int(d.tax_parent[x])
        key.append(tuple(reversed(path)))
    order = sorted(range(len(species)), key=lambda i: key[i])
    species = [species[i] for i in order]
    E = np.stack([node_seq[v] for v in species])          # (n_g, G) evolved, coordinate aligned
    n_g = cfg.n_genomes
    is_rev = (rng.random(n_g) < cfg.frac_rev_genomes) & (not cfg.protein)
    d.ref_names = ["G%05d" % g for g in range(n_g)]
    d.ref_seqs = [np.ascontiguousarray(revcomp(E[g]) if is_rev[g] else E[g]) for g in range(n_g)]
    d.ref_taxnode = np.array(species, dtype=np.uint32)

    K = min(cfg.n_cand, n_g)
    nb = G // block
    half = K // 2
    # cumulative block identities between genome g and g+off, off in [-half, K-half)
    offs = np.arange(-half, K - half)
    cum = np.zeros((n_g, K, nb + 1), dtype=np.int32)
    for oi, off in enumerate(offs):
        lo, hi = max(0, -off), min(n_g, n_g - off)
        if hi <= lo:
            continue
        eq = (E[lo:hi] == E[lo + off:hi + off]).reshape(hi - lo, nb, block).sum(axis=2)
        cum[lo:hi, oi, 1:] = np.cumsum(eq, axis=1)
    ref = FastReference()
    ref.block, ref.G, ref.E, ref.is_rev, ref.cum, ref.K, ref.half, ref.n_g = block, G, E, is_rev, cum, K, half, n_g
    return d, ref


def draw_query_lengths(cfg: SynthConfig, ref, rng, nq):
    """The first draw of generate_queries_fast on a fresh generator: the source-window lengths of nq queries."""
    Lq = (rng.integers(cfg.query_len[0], cfg.query_len[1] + 1, nq) // ref.block) * ref.block
    return np.clip(Lq, ref.block, ref.G)


def generate_queries_fast(cfg: SynthConfig, ref, rng, nq):
    """Second half of generate_fast: nq queries cut from the evolved genomes (+ noise) and their alignment
    records.  Returns (q_seqs: list of uint8 arrays, rec: dict of record columns, rec["q"] local to this call)."""
    alphabet = AA20 if cfg.protein else NT
    block, G, E, is_rev, cum, K, half, n_g = ref.block, ref.G, ref.E, ref.is_rev, ref.cum, ref.K, ref.half, ref.n_g
    Lq = draw_query_lengths(cfg, ref, rng, nq)
    g0 = rng.integers(0, n_g, nq)
    pblk = (rng.random(nq) * ((G - Lq) // block + 1)).astype(np.int64)
    p = pblk * block
    q_seqs = []
    for i0 in range(0, nq, 4096):
        i1 = min(nq, i0 + 4096)
        for i in range(i0, i1):
            src = E[g0[i], p[i]:p[i] + Lq[i]]
            q_seqs.append(src)
    # substitutions, chunked and vectorised over the concatenation
    lens = Lq.astype(np.int64)
    for i0 in range(0, nq, 8192):
        i1 = min(nq, i0 + 8192)
        cat = np.concatenate(q_seqs[i0:i1])
        cat = _mutate_subst(rng, cat, cfg.query_sub, alphabet)
        o = 0
        for i in range(i0, i1):
            seg = cat[o:o + lens[i]]
            o += lens[i]
            if cfg.query_indel > 0:
                seg = _mutate_indel(rng, seg, cfg.query_indel, alphabet)
            q_seqs[i] = np.ascontiguousarray(seg)
    qlen_final = np.array([len(s) for s in q_seqs], dtype=np.int64)

    # candidate genomes: window of K around g0, clipped to [0, n_g)
    lo = np.clip(g0 - half, 0, n_g - K)
    cand_g = lo[:, None] + np.arange(K)[None, :]                 # (nq, K)
    off_idx = cand_g - g0[:, None] + half                         # index into offs; may fall outside for clipped windows
    valid = (off_idx >= 0) & (off_idx < K)
    oi = np.clip(off_idx, 0, K - 1)
    b0 = pblk[:, None]
    b1 = b0 + (Lq // block)[:, None]
    ident = cum[g0[:, None], oi, b1] - cum[g0[:, None], oi, b0]   # (nq, K) exact block sums
    ident = np.where(valid, ident, (Lq[:, None] * 0.5).astype(np.int64))
    # what an aligner would report: matches shrink with the query's substitutions and deletions,
    # the alignment grows by the query's insertions (alnlen - identities ~ edit operations)
    ident = np.round(ident * (1.0 - cfg.query_sub) * (1.0 - cfg.query_indel / 2)).astype(np.int64)
    # partial records
    part = rng.random((nq, K)) < cfg.frac_partial
    a = (rng.random((nq, K)) * (Lq[:, None] // 4)).astype(np.int64) * part
    b = (rng.random((nq, K)) * (Lq[:, None] // 4)).astype(np.int64) * part
    qs = 1 + a
    qe = Lq[:, None] - b
    alnlen = qe - qs + 1
    ident = np.minimum((ident * alnlen) // Lq[:, None], alnlen - 1)
    if cfg.query_indel > 0:
        alnlen = np.round(alnlen * (1.0 + cfg.query_indel / 2)).astype(np.int64)
    # strictly decreasing identities per query (tie-free (score, identities))
    srt = np.argsort(-ident, axis=1, kind="stable")
    ident_s = np.take_along_axis(ident, srt, axis=1)
    dec = ident_s - np.arange(K)[None, :] * 0                     # enforce strict decrease
    for k in range(1, K):
        dec[:, k] = np.minimum(dec[:, k], dec[:, k - 1] - 1)
    ident = np.empty_like(ident)
    np.put_along_axis(ident, srt, dec, axis=1)
    keep = ident > 0
    fs = p[:, None] + qs
    fe = p[:, None] + qe
    rev = is_rev[cand_g]
    rstart = np.where(rev, G - fs + 1, fs)
    rstop = np.where(rev, G - fe + 1, fe)
    # rescale query coordinates when indels changed the query length
    scale = qlen_final[:, None] / Lq[:, None]
    qs2 = np.clip(np.round(qs * scale).astype(np.int64), 1, qlen_final[:, None])
    qe2 = np.clip(np.round(qe * scale).astype(np.int64), qs2, qlen_final[:, None])
    qs2 = np.where(scale == 1.0, qs, qs2)
    qe2 = np.where(scale == 1.0, qe, qe2)
    perm = np.argsort(rng.random((nq, K)), axis=1)                # file order is not score order
    def flat(x):
        return np.take_along_axis(x, perm, axis=1)[np.take_along_axis(keep, perm, axis=1)]
    qidx = np.broadcast_to(np.arange(nq)[:, None], (nq, K))
    ident_f = flat(ident)
    alnlen_f = flat(alnlen)
    rec = {
        "q": flat(qidx).astype(np.uint32), "qstart": flat(qs2).astype(np.uint32), "qstop": flat(qe2).astype(np.uint32),
        "r": flat(cand_g).astype(np.uint32), "rstart": flat(rstart).astype(np.uint32),
        "rstop": flat(rstop).astype(np.uint32),
        "score": (2.0 * ident_f - 3.0 * (alnlen_f - ident_f)).astype(np.float32),
        "ident": ident_f.astype(np.uint32), "alnlen": alnlen_f.astype(np.uint32),
    }
    return q_seqs, rec


def generate_fast(cfg: SynthConfig, block: int = 100) -> SynthData:
    """Vectorised generator for benchmark-sized nucleotide configs (same SynthData layout).

    Differences from generate(): query windows start on `block` boundaries, candidate identities are
    block-resolution ungapped identities of the source window (scaled by the query noise) -- they only
    need to be plausible and tie-free, not exact -- and genomes are not truncated.  Substitution-only
    queries (query_indel is ignored unless > 0, then a slower per-query path is used)."""
    rng = np.random.default_rng(cfg.seed)
    d, ref = build_reference_fast(cfg, rng, block)
    d.q_seqs, d.rec = generate_queries_fast(cfg, ref, rng, cfg.n_queries)
    d.q_names = ["Q%06d" % i for i in range(cfg.n_queries)]
    return d


# ---------------------------------------------------------------------------------------------------------
# Block-wise generation of very large batches (BASELINE.json configs[3] / [4] at full size): the batch is the
# concatenation of fixed-size query blocks, block k drawn from its own generator (seed, k), over ONE refpack /
# taxonomy drawn from the base seed.  The batch therefore does not depend on how many ranks generate it, a rank
# only generates the blocks that overlap its shard, and blocks can be generated by worker processes.
def block_rng(cfg: SynthConfig, k: int):
    return np.random.default_rng([int(cfg.seed), 7919, int(k)])


def block_query_lengths(cfg: SynthConfig, ref, n_blocks: int, block_queries: int):
    """Source-window lengths of every query of the batch, block by block (cheap: one draw per block) -- the
    work estimate n_cand * L^2 that cuts the batch into shards before anything else is generated."""
    return np.concatenate([draw_query_lengths(cfg, ref, block_rng(cfg, k), block_queries) for k in range(n_blocks)])


_BLOCK_STATE = {}


def _block_worker(k):
    cfg, ref, block_queries = _BLOCK_STATE["args"]
    q_seqs, rec = generate_queries_fast(cfg, ref, block_rng(cfg, k), block_queries)
    lens = np.array([len(s) for s in q_seqs], np.uint32)
    return k, np.concatenate(q_seqs), lens, rec


def generate_blocks(cfg: SynthConfig, d: SynthData, ref, block_ids, block_queries: int, workers: int = 1):
    """Blocks `block_ids` of the batch.  Returns (q_chars, q_len, segs, cands) with query ordinals / candidate
    offsets local to the returned range (block order).  workers > 1 forks worker processes -- call it before
    CUDA is initialised in this process."""
    block_ids = list(block_ids)
    _BLOCK_STATE["args"] = (cfg, ref, block_queries)
    if workers > 1 and len(block_ids) > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(min(workers, len(block_ids))) as pool:
            parts = pool.map(_block_worker, block_ids, chunksize=1)
    else:
        parts = [_block_worker(k) for k in block_ids]
    parts.sort(key=lambda x: x[0])
    if not parts:
        return np.zeros(0, np.uint8), np.zeros(0, np.uint32), np.zeros(0, SEG_DTYPE), np.zeros(0, CAND_DTYPE)
    # one output buffer, parts released as they are copied (a shard of the 2M-segment batch is tens of GB)
    chars = np.empty(sum(len(x[1]) for x in parts), np.uint8)
    lens, segs, cands = [], [], []
    qbase = cbase = cpos = 0
    for i in range(len(parts)):
        _, ch, ln, rec = parts[i]
        parts[i] = None
        chars[cpos:cpos + len(ch)] = ch
        cpos += len(ch)
        del ch
        tmp = SynthData(cfg=cfg)
        tmp.rec, tmp.ref_taxnode = rec, d.ref_taxnode
        sg, cd = segments_fast(tmp)
        sg["query_seq"] += qbase
        sg["cand_begin"] += cbase
        qbase += len(ln)
        cbase += len(cd)
        lens.append(ln); segs.append(sg); cands.append(cd)
    return chars, np.concatenate(lens), np.concatenate(segs), np.concatenate(cands)


def segments_fast(d: SynthData):
    """segments() for data where every query's records overlap into ONE segment (generate_fast with
    multi_segment_frac == 0): vectorised."""
    r = d.rec
    nrec = len(r["q"])
    order = np.lexsort((np.arange(nrec), r["qstop"], r["qstart"], r["q"]))
    q = r["q"][order]
    cand = np.zeros(nrec, dtype=CAND_DTYPE)
    cand["ref_seq"] = r["r"][order]; cand["rstart"] = r["rstart"][order]; cand["rstop"] = r["rstop"][order]
    cand["qstart"] = r["qstart"][order]; cand["qstop"] = r["qstop"][order]; cand["score"] = r["score"][order]
    cand["identities"] = r["ident"][order]; cand["alnlen"] = r["alnlen"][order]
    cand["node"] = d.ref_taxnode[r["r"][order]]
    starts = np.flatnonzero(np.concatenate([[True], q[1:] != q[:-1]]))
    counts = np.diff(np.concatenate([starts, [nrec]]))
    seg = np.zeros(len(starts), dtype=SEG_DTYPE)
    seg["query_seq"] = q[starts]; seg["cand_begin"] = starts; seg["cand_count"] = counts
    return seg, cand
