#!/usr/bin/env python
"""Print the SASS of one kernel (mangled-name substring) attributed to a source line range.

usage: sass_lines.py <lib.so> <mangled substring> <file> <first line> <last line>
"""
import glob, os, re, subprocess, sys, tempfile


def main():
    lib, kern, fname, l0, l1 = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    for cubin in glob.glob(tmp + "/*.cubin"):
        txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
        start = [i for i, l in enumerate(txt) if l.startswith(".text.") and kern in l]
        if not start:
            continue
        cur, n, tot = None, 0, 0
        for l in txt[start[0] + 1:]:
            if l.startswith(".text.") or l.startswith(".section"):
                break
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                cur = (m.group(1).split("/")[-1], int(m.group(2)))
                continue
            if re.search(r"/\*[0-9a-f]{4}\*/", l):
                tot += 1
                if cur and cur[0] == fname and l0 <= cur[1] <= l1:
                    print(cur[1], l.strip()[:120])
                    n += 1
        print(f"# {n} of {tot} instructions")
        return
    sys.exit("kernel not found")


if __name__ == "__main__":
    main()
