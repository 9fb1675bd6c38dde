"""Host-side mirror of the reference's `models/` package for the NeuS train-step hot path
(SURVEY.md §8b): same module / class / method names, constructor kwargs, parameter names and
`render()` dict, backed by the sm_100a kernels in ../csrc through the C ABI (include/fmov_b200.h).

To drive the reference's exp_runner.py unchanged, put this directory's parent on sys.path as
`models`."""
