"""Per-instruction view of an .ncu-rep (source page): SASS, executed count, stall samples and the top stall reasons.
    python scripts/ncu_source.py gpurun_out/prof_x.ncu-rep [min_samples]"""
import csv
import subprocess
import sys


def main(path, min_samples=0):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rows = list(csv.reader(lines[start:]))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    total = sum(int(r[ix["# Samples"]] or 0) for r in rows[1:] if len(r) == len(hdr))
    print(f"total samples {total}")
    for r in rows[1:]:
        if len(r) != len(hdr):
            continue
        n = int(r[ix["# Samples"]] or 0)
        if n < min_samples:
            continue
        st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
        st = " ".join(f"{c}:{v}" for v, c in st if v)
        print(f"{n:6d} {100.0 * n / max(total, 1):5.1f}%  x{r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip():70s} {st}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
