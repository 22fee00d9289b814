"""Device-side pieces of ``quantum_inferno.synth`` that sit next to the time-frequency path."""
