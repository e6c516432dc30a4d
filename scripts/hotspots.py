"""Join the SASS page of an ncu capture (per-instruction executed counts / stall samples) with `nvdisasm -g` line info of
the same cubin and print where the instructions of one kernel go, by source file:line group.
usage: python scripts/hotspots.py SRC_PAGE.csv DISASM.sass KERNEL_MANGLED_SUBSTRING [top]"""
import collections
import csv
import re
import sys


def load_disasm(path, kernel):
    """instruction index -> (file, line, inline-stack) for the .text section of `kernel`."""
    out = []
    cur_file, cur_line = "?", 0
    inside = False
    stack = ""
    for ln in open(path, errors="replace"):
        if ln.startswith(".text."):
            inside = kernel in ln
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur_file, cur_line = m.group(1).split("/")[-1], int(m.group(2))
            stack = m.group(3)
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            out.append((cur_file, cur_line, stack))
    return out


def main(srcpage, disasm, kernel, top=40):
    rows = list(csv.reader(open(srcpage)))
    hdr, data = rows[1], rows[2:]
    ie, isamp, ith = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
    info = load_disasm(disasm, kernel)
    print(f"ncu instructions {len(data)}, disasm instructions {len(info)}")
    n = min(len(data), len(info))
    by = collections.defaultdict(lambda: [0, 0, 0, 0])
    tot = [0, 0, 0]
    for i in range(n):
        e, s, t = int(data[i][ie]), int(data[i][isamp]), int(data[i][ith])
        k = (info[i][0], info[i][1])
        b = by[k]
        b[0] += e; b[1] += s; b[2] += t; b[3] += 1
        tot[0] += e; tot[1] += s; tot[2] += t
    byfile = collections.defaultdict(lambda: [0, 0, 0, 0])
    for (f, l), b in by.items():
        for q in range(4):
            byfile[f][q] += b[q]
    print("by file: exec% samples% lanes static-instr")
    for f, b in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
        print(f"  {f:28s} {100*b[0]/tot[0]:6.2f} {100*b[1]/max(tot[1],1):6.2f} {b[2]/max(b[0],1):5.1f} {b[3]:6d}")
    print("top lines: file:line exec% samples% lanes static-instr")
    for (f, l), b in sorted(by.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"  {f}:{l:<5d} {100*b[0]/tot[0]:6.2f} {100*b[1]/max(tot[1],1):6.2f} {b[2]/max(b[0],1):5.1f} {b[3]:5d}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 40)
