"""CPU oracle of the batched iLQR solve — test infrastructure only (see oracle/ilqr_oracle_impl.h)."""
