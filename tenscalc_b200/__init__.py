"""tenscalc_b200 -- batched, Blackwell-native backend for TensCalc's interior-point solvers."""
