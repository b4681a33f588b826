"""Static SASS opcode counts per kernel of an object / library: python tools/sass_opcodes.py skyrim_b200/csrc/graphcast_engine.o [filter]"""
import collections, re, subprocess, sys
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "UTMALDG", "HMMA", "LDGSTS", "STAS", "MUFU", "LDG", "STG", "LDL", "STL"]
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
flt = sys.argv[2] if len(sys.argv) > 2 else ""
cur, ops, n = None, collections.defaultdict(collections.Counter), collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r">\(.*", ">", cur).replace("void ", "").replace("sky::", "").replace("(int)", "")
        cur = re.sub(r"\(.*", "", cur) if "<" not in cur else cur
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        n[cur] += 1
        for k in KEYS:
            if m.group(1).startswith(k):
                ops[cur][k] += 1
print("| kernel | instrs | " + " | ".join(KEYS) + " |\n|---|---|" + "---|" * len(KEYS))
for k in sorted(ops):
    if flt in k:
        print(f"| `{k}` | {n[k]} | " + " | ".join(str(ops[k][x]) if ops[k][x] else "" for x in KEYS) + " |")
