"""Synthetic workload definitions for tests and bench.py (not product code): UNet skeletons with
the public SDXL / SD1.5 layer shapes and diffusers' class / attribute names, so that LyCORIS
presets select the same layers they would on the real models (diffusers is not installed here)."""
