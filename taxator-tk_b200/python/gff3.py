"""GFF3 rendering of prediction results, following PredictionRecord::print
(core/src/predictionrecord.hh:248-308) so that output can be compared byte for byte with
`taxator -a rpa`."""
import numpy as np

HEADER = "##gff-version 3\n"


def _fmt_float(v):
    # default ostream formatting of a float: 6 significant digits, %g style
    return "%g" % float(np.float32(v))


def tax_feature(lower, upper, support, parent, depth, taxid_of):
    """printFeatureTax: taxon_support_ has the same value at every rank (setNodeRange(l,u,s))."""
    out = []
    last = 0
    node = lower
    while node != upper:
        if support != last:
            out.append("%s:%d-" % (taxid_of(node), support))
            last = support
        node = int(parent[node])
    out.append(str(taxid_of(node)))
    if support != last:
        out.append(":%d" % support)
    return "".join(out)


def render(results, segs, q_names, q_lens, parent, depth, taxids, prev_ival=-1.0):
    """One line per segment, in segment order.  results: RESULT_DTYPE array."""
    lines = []
    ival_state = prev_ival  # PredictionRecord is reused: n==0 keeps the previous ival (taxator.cpp:66)
    for r, sg in zip(results, segs):
        q = int(sg["query_seq"])
        qlen = int(q_lens[q])
        if r["kind"] == 0:
            begin, end = 1, qlen
            ival = ival_state
        else:
            begin, end = int(r["qrstart"]), int(r["qrstop"])
            ival = float(r["ival"])
        ival_state = ival
        tax = tax_feature(int(r["lower_node"]), int(r["upper_node"]), int(r["support"]), parent, depth,
                          lambda n: taxids[n])
        s = "%s\ttaxator-tk\tsequence_feature\t%d\t%d\t%s\t.\t.\tseqlen=%d;tax=%s;rtax=%s" % (
            q_names[q], begin, end, _fmt_float(r["signal"]), qlen, tax, taxids[int(r["rtax_node"])])
        if 0.0 <= ival < 1.0:
            s += ";ival=" + _fmt_float(ival)
        lines.append(s + "\n")
    return lines
