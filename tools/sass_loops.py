"""Lists the backward-branch loops of one kernel in an object file with their instruction mix:
python tools/sass_loops.py atsc_b200/build/kernels.o k_polyILi1 [min_len]"""
import re, subprocess, sys, collections
obj, pat = sys.argv[1], sys.argv[2]
minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s+Function : ", txt)
for f in funcs:
    name = f.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = [(int(m.group(1), 16), m.group(2).strip()) for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", f)]
    print(name, len(ins), "instructions")
    addr = {a: k for k, (a, _) in enumerate(ins)}
    for k, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:[!A-Z0-9]+,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) <= a and int(m.group(1), 16) in addr:
            s = addr[int(m.group(1), 16)]
            body = ins[s:k + 1]
            if len(body) < minlen:
                continue
            ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", x).split()[0].split(".")[0] for _, x in body)
            f64 = sum(v for o, v in ops.items() if o in ("DFMA", "DMUL", "DADD", "DSETP"))
            print(f"  loop {ins[s][0]:#x}..{a:#x}: {len(body)} instr, f64 pipe {f64}, LDG {ops['LDG']}, LDL/STL {ops['LDL'] + ops['STL']}",
                  dict(ops.most_common(8)))
