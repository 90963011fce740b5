T=${1:-s4k}
timeout 1200 python -m pytest tests -m gpu -x -q -k "tcgen05_engine_vs_oracle or default_engine or moons or half" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log; tail -12 gpurun_out/${T}_pytest.log
python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);h=d['halfmoons'];print('halfmoons', h['ms'], h['engine'], h['e2e']['ms'], h['gpu_launches'])"
