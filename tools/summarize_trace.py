"""AMR_B200_TRACE time lines of a multi-rank bench run -> one table per C-ABI call of the LAST timed step: for every launch
the interval since the previous event of the same context (microseconds), per rank, with min / max over the ranks.
An interval covers the launch itself AND whatever it waited for (a neighbour's halo, the slowest rank of an all-reduce),
so a launch whose interval is long on every rank although it moves kilobytes is a synchronisation point.
Usage: python tools/summarize_trace.py gpurun_out/trace_n8full > profiles/r2_n8_timeline.txt"""
import glob
import re
import sys
from collections import defaultdict


def read(path):
    ev = []
    for line in open(path):
        m = re.match(r"\s*([0-9.]+)\s+([0-9.]+)\s+(.*)$", line.rstrip("\n"))
        if m:
            ev.append((float(m.group(1)), float(m.group(2)), m.group(3).strip()))
    return ev


def last_call(ev, name):
    """events of the last call `name` (up to the next amr_* entry)"""
    idx = [i for i, e in enumerate(ev) if e[2] == name]
    if not idx:
        return []
    i = idx[-1]
    out = []
    for e in ev[i + 1:]:
        if e[2].startswith("amr_"):
            break
        out.append(e)
    return out


def main(prefix):
    files = sorted(glob.glob(prefix + ".*"))
    by_rank = defaultdict(list)
    for f in files:
        rank = int(f[len(prefix) + 1:].split(".")[0])
        by_rank[rank].append(read(f))
    ranks = sorted(by_rank)
    print(f"{len(files)} trace files, ranks {ranks}")
    for call in ("amr_estimate", "amr_select_refine", "amr_map_fields_and_marking", "amr_select_unrefine"):
        rows = {}
        for r in ranks:
            best, best_key = [], (-1, -1)
            for ev in by_rank[r]:
                c = last_call(ev, call)
                # the contexts of the decomposed run carry halo launches; rank 0 also holds a serial context (set-up, checks)
                key = (1 if any("halo" in e[2] or "link" in e[2] for e in c) else 0, len(c))
                if c and key > best_key:
                    best, best_key = c, key
            rows[r] = best
        n = max(len(v) for v in rows.values())
        if n == 0:
            continue
        print(f"\n== {call}: last call of every rank, interval since the previous event [us]")
        print(f"{'launch':34s} " + " ".join(f"r{r:<6d}" for r in ranks) + "   min    max")
        tot = {r: 0.0 for r in ranks}
        for i in range(n):
            names = {rows[r][i][2] for r in ranks if i < len(rows[r])}
            vals = [rows[r][i][1] if i < len(rows[r]) else float("nan") for r in ranks]
            for r, v in zip(ranks, vals):
                if v == v:
                    tot[r] += v
            ok = [v for v in vals if v == v]
            print(f"{'/'.join(sorted(names))[:34]:34s} " + " ".join(f"{v:7.1f}" for v in vals) + f" {min(ok):6.1f} {max(ok):6.1f}")
        print(f"{'total':34s} " + " ".join(f"{tot[r]:7.1f}" for r in ranks))


if __name__ == "__main__":
    main(sys.argv[1])
