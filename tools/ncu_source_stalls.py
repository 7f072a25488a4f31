"""Per-kernel stall summary from an `ncu --set full --import-source on` report: stall-reason shares and the
instructions that collect the most samples.  Usage: python tools/ncu_source_stalls.py report.ncu-rep [out.txt]"""
import csv
import io
import subprocess
import sys


def main(rep, out=None):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    res = []
    i = 0
    seen = set()
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name":
            name = rows[i][1].replace("void ", "").split("(SellView")[0].split("(long")[0].split("(HaloPlan")[0]
            hdr = rows[i + 1]
            j = i + 2
            body = []
            while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
                if len(rows[j]) >= len(hdr) and rows[j][0].startswith("0x"):
                    body.append(rows[j])
                j += 1
            i = j
            if name in seen:
                continue
            seen.add(name)
            isrc, ismp = hdr.index("Source"), hdr.index("# Samples")
            stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
            idx = {h: hdr.index(h) for h in stalls}
            total = sum(int(r[ismp] or 0) for r in body)
            if not total:
                continue
            tot = {h: sum(int(r[idx[h]] or 0) for r in body) for h in stalls}
            res.append(f"== {name}: {total} samples")
            res.append("   " + ", ".join(f"{h[6:]} {v / total:.2f}" for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:5]))
            for r in sorted(body, key=lambda r: -int(r[ismp] or 0))[:6]:
                n = int(r[ismp] or 0)
                top = max(stalls, key=lambda h: int(r[idx[h]] or 0))
                res.append(f"   {n / total:5.2f}  {r[isrc].strip()[:70]:70s} {top[6:]}")
        else:
            i += 1
    s = "\n".join(res) + "\n"
    if out:
        open(out, "w").write(s)
    print(s)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
