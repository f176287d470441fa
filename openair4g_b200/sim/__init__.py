"""Link-level harness around the GPU decoder: what ulsim/dlsim do for the
turbo-coded shared channels, re-created on top of the C ABI because the reference simulators
cannot be built here.  Stimulus generation (TX chain, channel, demapper) is
plain numpy -- it is the test bench, not the product; decoding is always the CUDA library."""
