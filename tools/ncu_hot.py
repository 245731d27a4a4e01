"""Top stall sites of a kernel from an .ncu-rep captured with --import-source on (SASS view).
    python tools/ncu_hot.py gpurun_out/x.ncu-rep [N]"""
import csv
import io
import subprocess
import sys


def main(path, top=30):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    out = []
    kernel = None
    hdr = None
    for r in rows:
        if r and r[0] == "Kernel Name":
            kernel = r[1]
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr is None or len(r) < 6:
            continue
        out.append((int(r[2] or 0), int(r[5] or 0), r[1].strip(), kernel))
    total = sum(o[0] for o in out) or 1
    print(f"total samples {total}")
    idx = sorted(range(len(out)), key=lambda i: -out[i][0])[:top]
    for i in sorted(idx):
        s, ex, src, k = out[i]
        # show a little context: previous instruction
        print(f"{i:5d} {s:7d} {100.0 * s / total:5.1f}%  exec={ex:8d}  {src}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
