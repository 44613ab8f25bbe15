"""Host glue between the LyCORIS-shaped Python API and the sm_100a kernels (C-ABI via ctypes)."""
