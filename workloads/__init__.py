"""Checked, device-resident workload pipelines over the engine's C-ABI (BASELINE.json configs 3 and 5)."""
