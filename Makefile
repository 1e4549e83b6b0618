# Convenience targets (everything also works without make; see README.md)
PY ?= python

.PHONY: build test test-gpu bench bench-reference census ptxas reference clean

build:            ## compile gossipy_b200/_C*.so for sm_100a (nvcc cross-compiles without a GPU)
	$(PY) -c "import __graft_entry__ as g; g.build()"

test: build       ## CPU test-suite (multi-process paths over gloo + shared memory)
	$(PY) -m pytest tests/ -x -q -m "not gpu"

test-gpu: build   ## GPU test-suite (needs a B200)
	$(PY) -m pytest tests/ -x -q -m gpu

bench: build      ## headline benchmark on one GPU
	$(PY) bench.py --gpus 1

bench-reference: reference   ## the unmodified reference on the same configuration
	$(PY) bench.py --impl reference --gpus 1

census: build     ## SASS instruction census -> profiles/sass/CENSUS.md
	$(PY) tools/sass_census.py

ptxas:            ## registers / spills of every kernel -> profiles/PTXAS.md
	$(PY) -c "import __graft_entry__ as g; g.build(force=True)"
	$(PY) tools/ptxas_report.py

reference:        ## install the unmodified reference into baseline/_ref
	bash baseline/install_reference.sh

clean:
	rm -rf build gossipy_b200/_C*.so
