"""Helpers to compare prediction results with the committed golden GFF3 of the real reference."""
import json
import os

import oracle_lib as ol
import synth
import gff3

GOLDEN = os.path.join(ol.ROOT, "tests", "golden")
CASES = json.load(open(os.path.join(GOLDEN, "cases.json")))


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
