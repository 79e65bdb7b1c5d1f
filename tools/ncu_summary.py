#!/usr/bin/env python3
"""Condense ncu reports into the JSON committed under profiles/: python tools/ncu_summary.py out.json report1.ncu-rep [report2 ...]
(per kernel: duration, launch shape, pipe / issue utilisation incl. the fmaheavy / alu pipe counters, stall ratios, local-memory traffic,
DRAM bytes; for the Miller-loop and final-exponentiation kernels also the executed opcode mix with per-proof executed counts, the
EXECUTED IMAD.WIDE share of the multiplier issue rate and the stall-sample shares from the source page)."""
import collections, csv, io, json, re, subprocess, sys

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_fmalite.sum",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed.sum", "sm__cycles_active.avg", "launch__shared_mem_per_block_dynamic"]
HEAVY = ("k_miller", "k_miller_norm", "k_final_exp", "k_miller_lz", "k_final_exp_lz", "k_miller_norm_seg", "k_final_exp_stage", "k_lz", "k_lz2", "k_old", "k_pairing_lz")
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def ncu(rep, *args):
    return list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--csv"] + list(args), capture_output=True, text=True).stdout)))


def source_mix(rep, kernel):
    rows = ncu(rep, "--page", "source", "--kernel-name", kernel)
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address"); hdr = rows[h]; ix = {k: i for i, k in enumerate(hdr)}
    stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    ex, st, seen = collections.Counter(), collections.Counter(), set()
    for r in rows[h + 1:]:
        if len(r) < len(hdr) or r[0] in ("Address", "Kernel Name") or r[0] in seen:
            continue
        seen.add(r[0])
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]]); op = m.group(2) if m else "?"
        cls = "IMAD.WIDE" if op.startswith("IMAD.WIDE") else ("IMAD.other" if op.startswith("IMAD") else op.split(".")[0])
        ex[cls] += int(float(r[ix["Instructions Executed"]] or 0))
        for s in stalls:
            st[s[6:]] += int(float(r[ix[s]] or 0))
    T, S = sum(ex.values()), sum(st.values())
    return {"executed_warp_instructions": T, "executed_imad_wide_warp_instructions": ex["IMAD.WIDE"], "opcode_share": {k: round(v / T, 4) for k, v in ex.most_common(12)},
            "stall_sample_share": {k: round(v / S, 4) for k, v in st.most_common(8)}}


def main(out, reps):
    kernels = []
    for rep in reps:
        rows = ncu(rep, "--page", "raw")
        hdr, units = rows[0], rows[1]; ix = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = r[ix["Kernel Name"]].split("(")[0]
            d = {"kernel": name, "report": rep.split("/")[-1]}
            for w in WANT:
                if w in ix:
                    try:
                        d[w + (" [" + units[ix[w]] + "]" if units[ix[w]] else "")] = float(r[ix[w]].replace(",", ""))
                    except ValueError:
                        d[w] = r[ix[w]]
            for key, col in (("dram_bytes_read", "dram__bytes_read.sum"), ("dram_bytes_write", "dram__bytes_write.sum")):
                d[key] = float(r[ix[col]].replace(",", "")) * SCALE.get(units[ix[col]], 1)
            d["proofs"] = int(d.get("launch__grid_size", 0) * d.get("launch__block_size", 0))
            if name in HEAVY and d["proofs"] >= 1024:
                sp = d["source_page"] = source_mix(rep, name)
                # executed IMAD.WIDE per thread (= per proof for full blocks) and their share of the multiplier issue rate: one IMAD.WIDE per 4
                # cycles per scheduler, 4 schedulers per SM => peak = 1 warp-instruction per cycle per SM
                sp["executed_imad_wide_per_proof"] = round(sp["executed_imad_wide_warp_instructions"] * 32 / d["proofs"])
                sms, cyc = 148, d.get("sm__cycles_active.avg [cycle]")
                if cyc:
                    sp["executed_imad_wide_frac_of_issue_peak"] = round(sp["executed_imad_wide_warp_instructions"] / (sms * cyc), 4)
            kernels.append(d)
    json.dump({"source": "ncu --set full --metrics <fmaheavy / fmalite / alu pipe counters> --import-source on --clock-control none (recipes: tools/profile_r2.sh for bench.py --chunks 1 at 2^16 "
                         "RISC Zero-shape proofs, tools/r2_lzbench.sh for the layout microbenchmark); the .ncu-rep files are not kept",
               "note": "per-launch times under ncu are serialised and cold-cache: bench.py's live CUDA-event timings are the reported figures.  IMAD.WIDE issues once per 4 cycles "
                       "per scheduler, so a 25 % issue share of IMAD.WIDE would be 100 % of the integer-multiply roofline.",
               "kernels": kernels}, open(out, "w"), indent=1)
    for d in kernels:
        print(d["kernel"], d.get("gpu__time_duration.sum [ms]"), "ms  dram r/w GB", round(d["dram_bytes_read"] / 1e9, 3), round(d["dram_bytes_write"] / 1e9, 3), d.get("source_page", {}).get("opcode_share"))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
