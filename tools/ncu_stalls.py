"""Per-instruction stall summary of one kernel of an .ncu-rep (needs -lineinfo / --import-source on).
    python tools/ncu_stalls.py gpurun_out/prof.ncu-rep <kernel-name-substring> [instance] [top_n]"""
import csv, io, subprocess, sys

def sections(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    secs, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}; secs.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) >= len(cur["hdr"]) - 2:
            cur["rows"].append(r)
    return secs

def main():
    rep, pat = sys.argv[1], sys.argv[2]
    inst = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    secs = [s for s in sections(rep) if pat in s["name"]]
    s = secs[inst]
    hdr = s["hdr"]; col = {h: i for i, h in enumerate(hdr)}
    data = s["rows"]
    g = lambda r, h: int(float(r[col[h]] or 0))
    tot = sum(g(r, "# Samples") for r in data)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(g(r, h) for r in data) for h in stalls}
    print(s["name"], "| instances matching:", len(secs), "| samples", tot, "| SASS instructions", len(data))
    print("stall totals:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for r in sorted(data, key=lambda r: -g(r, "# Samples"))[:topn]:
        st = {h: g(r, h) for h in stalls}
        best = max(st.items(), key=lambda kv: kv[1])
        print(f"{g(r, '# Samples'):7d} {100.0 * g(r, '# Samples') / max(tot, 1):5.1f}% exec={g(r, 'Instructions Executed'):9d} {best[0][6:]:14s} {r[col['Source']].strip()[:110]}")

main()
