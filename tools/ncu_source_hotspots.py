"""Warp-state samples per SASS line from an `ncu --set full --import-source on` report.

    python tools/ncu_source_hotspots.py gpurun_out/r1_attn_bwd.ncu-rep [top_n] > profiles/..._source_hotspots.md

For each kernel in the report: total samples, the stall-reason mix, samples attributed to per-tile loops (grouped by the
execution count of the instruction, which identifies the warp role: softmax warps execute their loop 8x as often as the
single MMA / TMA warp), and the top-N instructions with their two main stall reasons.  This is the evidence behind the
round-2 candidate kernels (csrc/attention_r2.cu)."""
import collections
import csv
import subprocess
import sys


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    kern, cur = [], None
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kern.append(cur)
        elif r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and "hdr" in cur:
            cur["rows"].append(r)
    return kern


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    print(f"# Warp-state samples per SASS instruction — `{rep.split('/')[-1]}`\n")
    for k in load(rep):
        h = k["hdr"]
        ix = {n: i for i, n in enumerate(h)}
        stall = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        rows = k["rows"]
        tot = sum(int(r[ix["# Samples"]]) for r in rows)
        print(f"## `{k['name'][:110]}`\n")
        print(f"{len(rows)} SASS instructions, {tot} samples.\n")
        agg = collections.Counter({n: sum(int(r[ix[n]]) for r in rows) for n in stall})
        print("Stall mix: " + ", ".join(f"{n[6:]} {100 * v / tot:.0f} %" for n, v in agg.most_common(8) if v) + "\n")
        # group by execution count (role / loop identification)
        groups = collections.defaultdict(lambda: [0, 0, collections.Counter()])
        for r in rows:
            e = int(r[ix["Instructions Executed"]])
            key = e if e < 1000 else round(e, -3)
            g = groups[key]
            g[0] += 1
            g[1] += int(r[ix["# Samples"]])
            for n in stall:
                g[2][n] += int(r[ix[n]])
        print("| executions per instruction (≈) | instructions | samples | share | main stalls |\n|---|---|---|---|---|")
        for key, g in sorted(groups.items(), key=lambda kv: -kv[1][1])[:8]:
            ms = ", ".join(f"{n[6:]} {100 * v / max(g[1], 1):.0f} %" for n, v in g[2].most_common(3))
            print(f"| {key} | {g[0]} | {g[1]} | {100 * g[1] / tot:.0f} % | {ms} |")
        print(f"\nTop {top_n} instructions by samples:\n\n| # | SASS | samples | executed | stalls |\n|---|---|---|---|---|")
        top = sorted(enumerate(rows), key=lambda x: -int(x[1][ix["# Samples"]]))[:top_n]
        for i, r in sorted(top):
            st = sorted(((int(r[ix[n]]), n[6:]) for n in stall), reverse=True)[:2]
            src = r[ix["Source"]].strip().replace("|", "\\|")[:78]
            print(f"| {i} | `{src}` | {r[ix['# Samples']]} | {r[ix['Instructions Executed']]} | "
                  + ", ".join(f"{n} {v}" for v, n in st if v) + " |")
        print()


if __name__ == "__main__":
    main()
