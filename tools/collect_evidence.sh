#!/bin/bash
# One-GPU evidence run of a round: tests, bench lines, launch lists, ncu summaries. Everything lands in gpurun_out/.
# Usage (on the GPU box): bash tools/collect_evidence.sh
set -u
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/ev_pytest.txt
python bench.py 2>$O/ev_bench.err | tail -1 > $O/ev_bench_n1.json
python bench.py --impl reference 2>>$O/ev_bench.err | tail -1 > $O/ev_bench_reference_arm.json
python bench.py --workload config3 --no-extras --steps 2 --warmup 3 2>>$O/ev_bench.err | tail -1 > $O/ev_bench_config3_tc_auto.json
python bench.py --workload config3 --no-extras --steps 2 --warmup 3 --precision bf16x3 2>>$O/ev_bench.err | tail -1 > $O/ev_bench_config3_bf16x3.json
python bench.py --no-extras --precision bf16x3 2>>$O/ev_bench.err | tail -1 > $O/ev_bench_config2_bf16x3_same_box.json
python bench.py --no-extras --precision f16x2 2>>$O/ev_bench.err | tail -1 > $O/ev_bench_config2_f16x2_not_default.json
LRPCAP_TC_2SM=0 python bench.py --no-extras 2>>$O/ev_bench.err | tail -1 > $O/ev_bench_config2_no_pairs.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/ev_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-extras > /dev/null 2>&1
for P in h1f8 bf16x3; do
  PRECISION=$P timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_conv|last_dgrad" --csv --log-file $O/ev_launches_enc_$P.csv python tools/profile_encoder.py > /dev/null 2>&1
done
PRECISION=f16x2 RULE=presetA timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_conv|last_dgrad" --csv --log-file $O/ev_launches_enc_f16x2_preseta.csv python tools/profile_encoder.py > /dev/null 2>&1
PRECISION=h1f8 timeout 400 ncu --set full --clock-control none -k regex:"tc_conv|last_dgrad" --launch-skip 12 -c 13 -o $O/enc_eps python tools/profile_encoder.py > /dev/null 2>&1
python tools/ncu_summary.py $O/enc_eps.ncu-rep > $O/ev_ncu_enc_eps_h1f8_summary.csv; rm -f $O/enc_eps.ncu-rep
PRECISION=f16x2 RULE=presetA timeout 400 ncu --set full --clock-control none -k regex:"tc_conv|last_dgrad" --launch-skip 24 -c 13 -o $O/enc_pa python tools/profile_encoder.py > /dev/null 2>&1
python tools/ncu_summary.py $O/enc_pa.ncu-rep > $O/ev_ncu_enc_preseta_f16x2_summary.csv; rm -f $O/enc_pa.ncu-rep
N_IMAGES=64 WORDS_PER_IMAGE=1 timeout 400 ncu --set full --clock-control none -k regex:"tc_conv|simt_conv|pool_mask" -c 18 -o $O/fwd python tools/profile_encoder.py > /dev/null 2>&1
python tools/ncu_summary.py $O/fwd.ncu-rep > $O/ev_ncu_enc_forward64_summary.csv; rm -f $O/fwd.ncu-rep
python tools/bench_finetune.py --batch 64 > /dev/null 2>&1; cp $O/bench_finetune.json $O/ev_finetune_b64.json
python tools/bench_finetune.py --batch 256 > /dev/null 2>&1; cp $O/bench_finetune.json $O/ev_finetune_b256.json
cat $O/ev_pytest.txt
python - <<'PY'
import json
for f in ("ev_bench_n1", "ev_bench_reference_arm", "ev_bench_config3_tc_auto", "ev_bench_config3_bf16x3", "ev_bench_config2_bf16x3_same_box",
          "ev_bench_config2_f16x2_not_default", "ev_bench_config2_no_pairs", "ev_finetune_b64", "ev_finetune_b256"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, d.get("value", d.get("s_per_step")), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
