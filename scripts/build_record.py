"""Build provenance: compiler, flags, sha256 and the Blackwell SASS mnemonics of the in-tree library.
python scripts/build_record.py > profiles/r02_build_record.txt   (after `python -c "import __graft_entry__ as g; g.build()"`)"""
import hashlib, os, re, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
so = os.path.join(ROOT, "projects-2020-neural-svgd_b200", "nsvgd", "libnsvgd.so")
print("library:", os.path.relpath(so, ROOT), os.path.getsize(so), "bytes, sha256", hashlib.sha256(open(so, "rb").read()).hexdigest())
print("nvcc:", subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-2])
mk = open(os.path.join(ROOT, "projects-2020-neural-svgd_b200", "Makefile")).read()
print("flags:", re.search(r"NVCCFLAGS := (.*)", mk).group(1), "| ARCH", re.search(r"ARCH\s+:= (.*)", mk).group(1))
print("sources:", re.search(r"SRCS\s+:= (.*)", mk).group(1))
print("link:", [l.strip() for l in mk.splitlines() if "-shared -o $@ $(OBJS)" in l][0])
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
print("elf targets:", sorted(set(re.findall(r"arch = (sm_\w+)", sass))))
for mn in ("UTCHMMA.2CTA", "UTCHMMA", "UTMALDG.2D.2CTA", "UTMALDG.2D", "UTCBAR", "LDTM", "STTM", "MUFU.EX2", "UCGABAR_ARV", "SYNCS.ARRIVE"):
    print(f"SASS {mn}: {len(re.findall(r'\b' + re.escape(mn) + r'\b', sass))}")
syms = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
print("exported nsvgd_* symbols:", len([l for l in syms.splitlines() if " T nsvgd_" in l]))
print("links cuBLAS:", "libcublas" in subprocess.run(["ldd", so], capture_output=True, text=True).stdout)
print("build mode: in-tree `make` driven by __graft_entry__.build(); the .so is git-ignored and travels to the GPU box with the gpurun snapshot")
