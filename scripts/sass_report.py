"""SASS evidence: per-kernel mnemonic histogram of build/libdtb200.so (cuobjdump -sass), written to profiles/sass/.
Proves the Blackwell-native paths: UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTMALDG/UTMASTG/UTMAREDG (TMA),
UTCBAR (tcgen05.commit), SYNCS (mbarrier), LDG/STG/RED on peer pointers in the exchange kernels, LDGMC/STGMC (multimem.ld_reduce /
multimem.st through the NVSwitch), FFMA2/FMUL2/FADD2 (packed fp32x2 epilogue math)."""
import collections, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "distributedtraining_b200", "build", "libdtb200.so")
out_dir = os.path.join(ROOT, "profiles", "sass")
os.makedirs(out_dir, exist_ok=True)
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = re.compile(r"^(UTC\w*|UTMA\w*|UBLKCP|LDTM|STTM|SYNCS|HMMA|LDGSTS|RED|ATOM\w*|LDGMC|STGMC|REDGMC|LDG|STG|LDS|STS|MUFU|BAR|UCGABAR\w*|MEMBAR|ERRBAR|CCTL|UTCBAR|ACQBULK|ELECT|FFMA2|FMUL2|FADD2)")
kernels, cur, name = {}, None, None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        name = re.sub(r"\(.*", "", name)
        cur = kernels.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["_total"] += 1
        base = op.split(".")[0]
        if KEY.match(base):
            cur[op if base.startswith(("UTC", "UTMA", "LDTM", "SYNCS")) else base] += 1
summary = {}
for k, c in sorted(kernels.items()):
    short = re.sub(r"^void dtb::|^dtb::", "", k)
    summary[short] = {"instructions": c["_total"], **{op: n for op, n in sorted(c.items()) if op != "_total"}}
json.dump(summary, open(os.path.join(out_dir, "mnemonics_by_kernel.json"), "w"), indent=1)
with open(os.path.join(out_dir, "README.md"), "w") as f:
    f.write("# SASS evidence (cuobjdump -sass build/libdtb200.so, sm_100a)\n\n| kernel | instrs | tcgen05.mma (UTC*MMA) | tcgen05.ld (LDTM) | TMA (UTMALDG/STG/REDG) | mbarrier (SYNCS) | multimem (LDGMC/STGMC) | fp32x2 (FFMA2/FMUL2/FADD2) | HMMA (legacy) |\n|---|---|---|---|---|---|---|---|---|\n")
    for k, c in summary.items():
        g = lambda pre: sum(v for op, v in c.items() if op.startswith(pre))
        f.write(f"| `{k[:90]}` | {c['instructions']} | {g('UTC') - g('UTCBAR') - g('UTCATOM')} | {g('LDTM')} | {g('UTMA')} | {g('SYNCS')} | {g('LDGMC') + g('STGMC') + g('REDGMC')} | {g('FFMA2') + g('FMUL2') + g('FADD2')} | {g('HMMA')} |\n")
print(json.dumps({k: v for k, v in list(summary.items())[:3]}, indent=0)[:600])
print("kernels:", len(summary))

# ---- per-kernel LISTINGS of the hot loops (a histogram is not a listing) ----
LIST = [("sm100_gemm_kernel<false, false, false, 2, false>", "gemm_bf16_2sm_kmajor"), ("sm100_gemm_kernel<true, true, true, 2, false>", "gemm_wgrad_f32_2sm"),
        ("sm100_gemm_kernel<false, false, false, 2, true>", "gemm_fp8_2sm"),
        ("attn_fwd_small_kernel", "attn_fwd_small"), ("attn_bwd_small_kernel", "attn_bwd_small"), ("attn_fwd_kernel", "attn_fwd_tiled"),
        ("gather_avg_kernel<0>", "gather_avg_fp32"), ("seg_dot_kernel<0>", "seg_dot_fp32"), ("shard_transpose_kernel<0>", "shard_transpose_fp32"),
        ("nvls_avg_kernel", "nvls_avg"), ("shard_pull_reset_kernel", "shard_pull_reset"), ("adamw_kernel<0>", "adamw")]
HOT = re.compile(r"UTC\w*MMA|UTMALDG|UTMASTG|UTMAREDG|LDTM|UTCBAR|SYNCS|LDGMC|STG\.E\.128\.STRONG\.SYS|MULTIMEM|ELECT")
blocks, cur_name, cur_lines = {}, None, []
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        if cur_name:
            blocks[cur_name] = cur_lines
        cur_name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        cur_lines = []
        continue
    if cur_name and re.match(r"\s*/\*[0-9a-f]{4}\*/", line):
        cur_lines.append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line.rstrip()))
if cur_name:
    blocks[cur_name] = cur_lines
index = []
for pat, short in LIST:
    hit = [k for k in blocks if pat in k]
    if not hit:
        continue
    k = sorted(hit, key=len)[0]
    lines = blocks[k]
    hot = [i for i, l in enumerate(lines) if HOT.search(l)]
    if len(lines) > 900 and hot:  # long kernel: keep the regions around the tensor-core / TMA / multicast instructions
        keep = set()
        for i in hot:
            keep.update(range(max(0, i - 6), min(len(lines), i + 7)))
        sel, prev = [], -2
        for i in sorted(keep):
            if i != prev + 1:
                sel.append("        ...")
            sel.append(lines[i])
            prev = i
    else:
        sel = lines
    with open(os.path.join(out_dir, short + ".sass"), "w") as f:
        f.write(f"// {k}\n// cuobjdump -sass build/libdtb200.so (sm_100a); {len(lines)} instructions" + (", excerpt around UTC*MMA / UTMA* / LDTM / SYNCS / multicast\n" if len(sel) != len(lines) else "\n"))
        f.write("\n".join(sel) + "\n")
    index.append((short, k, len(lines), len(hot)))
with open(os.path.join(out_dir, "README.md"), "a") as f:
    f.write("\n## Listings\n\n`multimem.st` is emitted as `STG.E.128.STRONG.SYS` to the MULTICAST virtual address (only `multimem.ld_reduce` has its own opcode, `LDGMC`); "
            "the address is what makes it a switch-replicated store.\n\n| file | kernel | instructions | tensor-core / TMA / TMEM / barrier instructions |\n|---|---|---|---|\n")
    for short, k, n, h in index:
        f.write(f"| `{short}.sass` | `{k[:100]}` | {n} | {h} |\n")
print("listings:", [i[0] for i in index])
