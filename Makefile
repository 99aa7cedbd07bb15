# Convenience targets (the driver uses __graft_entry__.py / pytest / bench.py directly).
PY ?= python
GPURUN ?= /usr/local/graft/bin/gpurun

build:            ## nvcc -gencode arch=compute_100a,code=sm_100a -> transformer-mm-explainability_b200/libmmx.so
	$(PY) __graft_entry__.py

test-cpu:         ## oracle vs reference goldens, ABI symbols, gloo sharding logic
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu:         ## parity through the C ABI (needs a B200)
	$(PY) -m pytest tests -q -m gpu

golden:           ## regenerate tests/golden/*.npz from the unmodified reference (build container only: needs /root/reference)
	$(PY) -m oracle.make_golden

bench:
	$(PY) bench.py --gpus 1

sanitize:         ## compute-sanitizer over every kernel family (small shapes)
	compute-sanitizer --tool memcheck --error-exitcode 3 $(PY) profiles/sanitize_run.py
	compute-sanitizer --tool racecheck --error-exitcode 3 $(PY) profiles/sanitize_run.py

launches:         ## ncu launch list of two interpret() steps -> markdown table
	ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 800 --csv --log-file gpurun_out/launches.csv $(PY) bench.py --steps 1 --warmup 1 --no-cpu-baseline
	$(PY) profiles/summarize_launches.py gpurun_out/launches.csv

.PHONY: build test-cpu test-gpu golden bench sanitize launches
