# Convenience targets; the driver uses __graft_entry__.build() / pytest / bench.py directly.
PY ?= python

build:            ## libhark.so (nvcc, sm_100a) + the CPU oracle
	$(PY) -c "import __graft_entry__ as g; g.build()"

test-cpu: build   ## oracle vs golden vectors, host logic, gloo world-1/2/3 sharding, bench contract
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu:         ## parity of every operator through the C-ABI (needs a B200)
	$(PY) -m pytest tests -q -m gpu

bench:            ## config 2: rows/s, roofline, e2e, cpu_baseline (one JSON line)
	$(PY) bench.py

sanitize:         ## compute-sanitizer memcheck + racecheck over every kernel family at small shapes (needs a B200)
	compute-sanitizer --tool memcheck --error-exitcode 3 $(PY) -m pytest tests/test_gpu_sanitize_shapes.py -m gpu -q -x -p no:cacheprovider
	compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 3 $(PY) -m pytest tests/test_gpu_sanitize_shapes.py -m gpu -q -x -p no:cacheprovider

clean:
	$(MAKE) -C harkdb_b200/csrc clean
	$(MAKE) -C oracle clean

.PHONY: build test-cpu test-gpu bench sanitize clean
