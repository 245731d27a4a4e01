"""Stall samples per CUDA source line of one kernel, from an .ncu-rep captured with --import-source on.
    python tools/ncu_lines.py gpurun_out/x.ncu-rep kernel_regex [N]"""
import csv
import io
import subprocess
import sys


def main(path, kern, top=30):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kern],
                         capture_output=True, text=True).stdout
    cur, out, seen_fn = None, [], 0
    for r in csv.reader(io.StringIO(raw)):
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) >= 2 and r[0] == "Function Name":
            seen_fn += 1
            if seen_fn > 1 and False:
                break
        elif len(r) > 8 and r[0].isdigit():
            num = lambda v: int(v) if v.strip().lstrip("-").isdigit() else 0
            out.append((num(r[4]), cur, int(r[0]), r[1].strip()[:120], num(r[7])))
    agg = {}
    for s, f, l, src, ex in out:
        a = agg.setdefault((f, l), [0, src, 0])
        a[0] += s
        a[2] += ex
    tot = sum(a[0] for a in agg.values()) or 1
    print(f"total samples {tot}")
    for (f, l), (s, src, ex) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{s:7d} {100.0 * s / tot:5.1f}%  {f}:{l}  exec={ex}  | {src}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
