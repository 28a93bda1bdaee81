# Convenience targets (the driver uses __graft_entry__.build()).  GPU targets need a B200: run them through gpurun.
.PHONY: build test sanitize configs
build:
	python -c "import __graft_entry__ as g; g.build()"
test:
	python -m pytest tests -x -q -m "not gpu"
# compute-sanitizer memcheck / synccheck / racecheck over the tcgen05 / TMA / cluster kernels at small shapes (profiles/sanitize_r02.log)
sanitize:
	bash tools/sanitize.sh
# bench lines of every BASELINE.json configuration (profiles/bench_*_r02.json)
configs:
	bash tools/run_configs.sh
