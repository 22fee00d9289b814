"""Host-side helpers of ``quantum_inferno.utilities`` that the time-frequency hot path touches."""
