"""profiles/r<round>_summary.md from the ncu launch list (CSV) and the raw page of the full capture (CSV).
usage: python tools/make_profile_summary.py profiles/r2_launches_bench.csv profiles/r2_ncu_full.csv "Round 2" > profiles/r2_summary.md
(raw page: ncu -i gpurun_out/r2_full.ncu-rep --page raw --csv > profiles/r2_ncu_full.csv)"""
import collections
import csv
import sys

OUTSIDE = ("k_posegraph", "k_lcd_store_row", "k_lcd_score", "k_lcd_score_umma", "k_lcd_q2h", "k_calc_conv", "k_calc_conv_umma", "k_calc_nhwc_split", "k_calc_resize", "k_calc_blur", "k_calc_pool", "k_calc_lrn",
           "k_calc_norm", "k_calc_normalize", "k_pack_record", "k_unpack_records")
TITLE = sys.argv[3] if len(sys.argv) > 3 else "Round 1, final kernels (session 4)"
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault(r[4].split("(")[0].replace("void ", "").split("<")[0], [0, 0.0])
    a[0] += 1
    a[1] += float(r[14])
ours = {k: v for k, v in agg.items() if k.startswith("k_")}
tot = sum(v[1] for v in ours.values())
step = sum(v[1] for k, v in ours.items() if k not in OUTSIDE)
print(f"# {TITLE} — ncu evidence (B200, 1 GPU)\n")
print(f"Launch list: `ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline`")
print(f"(raw CSV: `{sys.argv[1].split('/')[-1]}`; 64 stereo pairs = 128 images + 64 BA windows per step). Times are cold-cache and serialised: compare shares.")
print("`k_posegraph` (a config-4 extra outside the timed region) is listed for reference; the last column is the share among the kernels of the timed step.\n")
print("| kernel | launches | avg us | share of our kernels | share of the step's kernels |\n|---|---|---|---|---|")
for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1]):
    inside = f"{100 * v[1] / step:.1f}%" if k not in OUTSIDE else "—"
    print(f"| {k} | {v[0]} | {v[1] / v[0] / 1e3:.1f} | {100 * v[1] / tot:.1f}% | {inside} |")
r = list(csv.reader(open(sys.argv[2])))
h = r[0]
units = r[1]
TO_MS = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}
TO_MB = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
cols = [("gpu__time_duration.sum", "ms"), ("dram__bytes_read.sum", "MB rd"), ("dram__bytes_write.sum", "MB wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"), ("launch__registers_per_thread", "regs"),
        ("smsp__inst_executed.sum", "warp inst")]
print(f"\nFull capture (`ncu --set full --clock-control none --import-source on -k regex:^k_(...) -s 48 -c 16`, one whole step; CSV of the raw page: "
      f"`{sys.argv[2].split('/')[-1]}`), per launch (128 images / 64 matching problems / 64 BA windows):\n")
print("| kernel | " + " | ".join(c[1] for c in cols) + " |\n|" + "---|" * (len(cols) + 1))
for row in r[2:]:
    vals = []
    for c, _ in cols:
        v = row[h.index(c)] if c in h else ""
        try:
            u = units[h.index(c)] if c in h else ""
            scale = TO_MS.get(u, 1.0) if c.startswith("gpu__time") else TO_MB.get(u, 1.0) if c.startswith("dram__bytes") else 1.0
            v = f"{float(v) * scale:.4g}"
        except ValueError:
            pass
        vals.append(v)
    print("| " + row[h.index("Kernel Name")].split("(")[0] + " | " + " | ".join(vals) + " |")
print("\nSASS evidence (cuobjdump -sass libslamb200.so): `UTCIMMA` (tcgen05.mma kind::i8), `UTMALDG.2D` (TMA), `LDTM.x32` (tcgen05.ld), "
      "`UTCBAR` (tcgen05.commit) in `k_hamming_umma`; `UTMALDG.3D` in `k_fast_cells`, `k_blur` and `k_describe` (per-keypoint patch boxes); `ACQBULK` (griddepcontrol.wait) and `PREEXIT` (griddepcontrol.launch_dependents) at the top of every extract / match kernel (programmatic dependent launch).")
