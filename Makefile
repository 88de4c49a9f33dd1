# Convenience targets (the driver uses __graft_entry__.build(); these do the same from a shell).
.PHONY: all lib oracle emu test-cpu clean
all: lib oracle
lib:            ## libnemo_fct.so, sm_100a (nvcc cross-compiles without a GPU)
	$(MAKE) -C nemo-fmi-devel_b200/csrc -j4
oracle:         ## CPU oracle (test infrastructure)
	$(MAKE) -C oracle liboracle.so
emu:            ## host emulation of the column kernels (test infrastructure)
	$(MAKE) -C tests/emu libemu.so
test-cpu: all
	python -m pytest tests -q -m "not gpu"
clean:
	$(MAKE) -C nemo-fmi-devel_b200/csrc clean
	$(MAKE) -C oracle clean
	$(MAKE) -C tests/emu clean
