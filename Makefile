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

clean:
	$(MAKE) -C harkdb_b200/csrc clean
	$(MAKE) -C oracle clean

.PHONY: build test-cpu test-gpu bench clean
