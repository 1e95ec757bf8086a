"""Per-CUDA-source-line instruction histogram of one ncu report (ncu -i REP --page source --print-source cuda,sass --csv):
share of the kernel's warp-level instructions and active lanes per instruction, for the hottest source lines."""
import csv, subprocess, sys

def hist(rep, top=40):
    txt = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(txt.splitlines()))
    agg, tot, f = [], 0, "?"
    for r in rows:
        if r and r[0] == "File Path":
            f = r[1].split("/")[-1]
        elif len(r) > 9 and r[0].isdigit():
            try:
                ie, te = int(r[7]), int(r[8])
            except ValueError:
                continue
            if ie:
                agg.append((ie, te, f, int(r[0]), r[1].strip()))
                tot += ie
    out = [f"total warp-level instructions: {tot}", "share%  lanes  file:line  source"]
    for ie, te, f, ln, src in sorted(agg, reverse=True)[:top]:
        out.append(f"{100 * ie / tot:6.2f}  {te / ie:5.1f}  {f}:{ln}  {src[:120]}")
    return "\n".join(out)

if __name__ == "__main__":
    print(hist(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40))
