"""First n queries (and their records) of a SynthData, same refpack / taxonomy (shared by tests)."""
import synth


def subset(d, n_queries):
    s = synth.SynthData(cfg=d.cfg)
    s.tax_ids, s.tax_parent, s.tax_rank = d.tax_ids, d.tax_parent, d.tax_rank
    s.ref_names, s.ref_seqs, s.ref_taxnode = d.ref_names, d.ref_seqs, d.ref_taxnode
    s.q_names, s.q_seqs = d.q_names[:n_queries], d.q_seqs[:n_queries]
    m = d.rec["q"] < n_queries
    s.rec = {k: v[m] for k, v in d.rec.items()}
    return s
