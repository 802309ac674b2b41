"""Row comparison helpers shared by the parity tests (TEST INFRASTRUCTURE)."""


def _norm(key_type, row):
    """DUP / TRA RNAMES come from Python set iteration in the reference (PYTHONHASHSEED-dependent,
    resolveDUP.py:82,96; resolveTRA.py:182): compare them as sets."""
    if key_type in ("DUP", "TRA"):
        row = list(row)
        row[-1] = ",".join(sorted(row[-1].split(",")))
    return [str(x) for x in row]


def diff_rows(ref, got, max_report=5):
    """ref/got: {(type, chrom): rows}.  Returns list of human-readable differences."""
    msgs = []
    keys = sorted(set(ref) | set(got))
    for k in keys:
        a = [_norm(k[0], r) for r in ref.get(k, [])]
        b = [_norm(k[0], r) for r in got.get(k, [])]
        if a == b:
            continue
        if len(a) != len(b):
            msgs.append("%s: %d reference rows vs %d" % (k, len(a), len(b)))
        for i, (x, y) in enumerate(zip(a, b)):
            if x != y:
                msgs.append("%s row %d:\n  ref %s\n  got %s" % (k, i, x, y))
                if len(msgs) >= max_report:
                    return msgs
    return msgs
