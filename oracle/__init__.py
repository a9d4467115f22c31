"""CPU oracle of the executor path - TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this
package, and only as the checker or the timed CPU baseline.  The product (explainable-spatial-vqa_b200/)
never imports it and has no CPU fallback.
"""
