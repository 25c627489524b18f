"""lav_b200 test suite: `-m "not gpu"` (oracle vs golden, ABI, host logic, 2-rank gloo) and `-m gpu` (parity through the C ABI)."""
