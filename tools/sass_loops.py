#!/usr/bin/env python
"""List the innermost loops of one kernel of a .so with instruction mix and the RF-cycle model.
    python tools/sass_loops.py lib.so '<mangled kernel name>' [min_instructions]"""
import re
import subprocess
import sys
import tempfile

sys.path.insert(0, __import__("os").path.dirname(__file__))


def main():
    lib, fn = sys.argv[1], sys.argv[2]
    min_len = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", fn, lib], capture_output=True, text=True).stdout
    tmp = tempfile.NamedTemporaryFile("w", suffix=".sass", delete=False)
    tmp.write(txt)
    tmp.close()
    loops = []
    for line in txt.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4})\*/\s+(.*?);", line)
        if not m:
            continue
        a = int(m.group(1), 16)
        b = re.search(r"BRA\s+(0x[0-9a-f]+)", m.group(2))
        if b and int(b.group(1), 16) < a:
            loops.append((int(b.group(1), 16), a))
    # innermost = loops that contain no other loop
    inner = [l for l in loops if not any(o != l and o[0] >= l[0] and o[1] <= l[1] for o in loops)]
    for lo, hi in inner:
        if (hi - lo) // 16 < min_len:
            continue
        body = [l for l in txt.splitlines() if (mm := re.match(r"\s*/\*([0-9a-f]{4})\*/", l)) and lo <= int(mm.group(1), 16) <= hi]
        tags = [t for t in ("FSEL", "SHFL", "DFMA", "STS", "FMNMX") if any(t in l for l in body)]
        print(f"loop {lo:#x}..{hi:#x} ({(hi - lo) // 16 + 1} instr) {tags}")
        subprocess.run([sys.executable, __import__("os").path.join(__import__("os").path.dirname(__file__), "sass_rf_model.py"),
                        tmp.name, hex(lo), hex(hi)])


if __name__ == "__main__":
    main()
