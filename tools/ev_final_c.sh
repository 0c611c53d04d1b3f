# final single-GPU evidence, part C: smoke, full tests, general fuzz, whole-program timing with a held context
set -x
O=gpurun_out/fc
mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python tools/fuzz_gpu.py 90 21 > $O/fuzz.log 2>&1; echo "rc=$?" >> $O/fuzz.log
QUICK=1 GENOME=20000000 READS=2000000 timeout 1500 python tools/dropin_speed.py > $O/dropin_speed_2M_t16.json 2> $O/dropin_speed_2M_t16.err
