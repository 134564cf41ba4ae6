"""usage: python tools/ptxas_regs.py file.cu [filter]  -- registers / spills / smem per kernel (nvcc -Xptxas -v, sm_100a)"""
import re, subprocess, sys
src = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--expt-relaxed-constexpr",
                      "-Xptxas", "-v", "-c", src, "-o", "/dev/null"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
name = None
spill = ""
for line in out.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name)
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores", line)
    if m:
        spill = f"stack {m.group(1)} spill {m.group(2)}"
    m = re.search(r"Used (\d+) registers(.*)", line)
    if m and name and flt in name:
        print(f"{int(m.group(1)):4d} regs  {spill:22s} {name}  {m.group(2).strip()[:60]}")
if "error" in out:
    print(out)
