#!/bin/bash
# First GPU call of the next round (about 2 GPU-minutes): the unmeasured candidates left by round 1.
#   gpurun --timeout 300 -- 'bash tools/round2_first_call.sh'
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
o=gpurun_out/r2_first
( LDIFF_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -q -m gpu ) > ${o}_experimental_tests.log 2>&1
python tools/kbench.py > ${o}_kbench_default.txt 2>&1
LDIFF_DT_TMA=1 timeout 120 python tools/kbench.py > ${o}_kbench_tma.txt 2>&1
for cfg in "" "LDIFF_DT_TMA=1" "LDIFF_PASS_ZERO_IN_CHAINS=1" "LDIFF_DT_TMA=1 LDIFF_PASS_ZERO_IN_CHAINS=1"; do
  env $cfg TAG="[$cfg]" timeout 120 python tools/pass_sched.py >> ${o}_pass.txt 2>&1
done
LDIFF_DT_TMA=1 timeout 120 python tools/pass_persist.py "(0, 0, (-2, -1, 0, 0, -2), False)" "(52, 0, (-2, -1, -3, 0, -2), True)" "(52, 96, (-2, -1, -3, 0, -2), True)" > ${o}_pass_persist_tma.txt 2>&1
tail -n 3 ${o}_experimental_tests.log
grep -h decode_tail ${o}_kbench_default.txt ${o}_kbench_tma.txt
cat ${o}_pass.txt ${o}_pass_persist_tma.txt
