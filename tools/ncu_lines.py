"""Summarise an ncu report per CUDA source line: `python tools/ncu_lines.py <report.ncu-rep> [top_n]`.
Uses `ncu --page source --csv --print-source cuda,sass` (needs -lineinfo and --import-source on)."""
import csv
import subprocess
import sys


def main(path, top_n=30):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, hdr = None, None
    lines = []  # (file, line, source, samples, inst, thread_inst)
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file, hdr = r[1], None
            continue
        if r and r[0] == "Line No":
            hdr = r
            i_s, i_i, i_t = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
            i_l = hdr.index("stall_long_sb") if "stall_long_sb" in hdr else None
            continue
        if hdr is None or not r or r[0] == "" or len(r) <= i_t:
            continue
        try:
            lines.append((cur_file.rsplit("/", 1)[-1], int(r[0]), r[1], float(r[i_s] or 0), float(r[i_i] or 0), float(r[i_t] or 0),
                          float(r[i_l] or 0) if i_l else 0.0))
        except ValueError:
            pass
    ts = sum(x[3] for x in lines) or 1
    ti = sum(x[4] for x in lines) or 1
    print(f"total samples {ts:.0f}  total warp-inst {ti:.0f}")
    for f, ln, src, s, i, t, l in sorted(lines, key=lambda x: -x[3])[:top_n]:
        print(f"{f}:{ln:<4d} smp {100*s/ts:5.1f}%  inst {100*i/ti:5.1f}%  thr/inst {t/max(i,1):5.1f}  long_sb {100*l/ts:5.1f}%  | {src.strip()[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)


def regions(path, spec):
    """spec: list of (name, file, lo, hi) -> share of samples / instructions per region"""
    import io, contextlib
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, hdr = None, None
    acc = {n: [0.0, 0.0, 0.0] for n, *_ in spec}
    acc["other"] = [0.0, 0.0, 0.0]
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file, hdr = r[1].rsplit("/", 1)[-1], None
            continue
        if r and r[0] == "Line No":
            hdr = r
            i_s, i_i, i_t = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
            continue
        if hdr is None or not r or r[0] == "" or len(r) <= i_t:
            continue
        try:
            ln = int(r[0]); s = float(r[i_s] or 0); i = float(r[i_i] or 0); t = float(r[i_t] or 0)
        except ValueError:
            continue
        for n, f, lo, hi in spec:
            if cur_file == f and lo <= ln <= hi:
                a = acc[n]; break
        else:
            a = acc["other"]
        a[0] += s; a[1] += i; a[2] += t
    ts = sum(a[0] for a in acc.values()) or 1
    ti = sum(a[1] for a in acc.values()) or 1
    for n, a in acc.items():
        print(f"{n:12s} samples {100*a[0]/ts:5.1f}%  warp-inst {100*a[1]/ti:5.1f}% ({a[1]:.0f})  thr/inst {a[2]/max(a[1],1):5.1f}")
