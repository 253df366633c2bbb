"""Roofline denominators measured live (debug tool, not a pytest file)."""
import sys
sys.path.insert(0, ".")
from jaxili_b200.engine import fp32_peak_tflops, tf32_peak_tflops
print("fp32:", fp32_peak_tflops(0))
for it in (2000, 20000, 100000):
    print("tf32 iters", it, tf32_peak_tflops(0, it))
