#!/usr/bin/env python
"""The fuzz test of the full-semantics kernels (tests/test_gpu_flex.py: random configurations and masked command sequences vs the
oracle, bitwise across launch splits) over many more seeds than the test suite runs.  usage: python tools/flex_fuzz_many.py [first] [count]"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cdpr_simulation_b200 as cb
import test_gpu_flex as t
first = int(sys.argv[1]) if len(sys.argv) > 1 else 6
count = int(sys.argv[2]) if len(sys.argv) > 2 else 100
kinds = collections.Counter()
orig = cb.CdprBatch.set_independent
def spy(self, on=True):
    r = orig(self, on); kinds[self.kernel_detail] += 1; return r
cb.CdprBatch.set_independent = spy
bad = []
for seed in range(first, first + count):
    try:
        t.test_flex_random_command_sequences_against_the_oracle(None, seed)
    except AssertionError as e:
        bad.append((seed, str(e)[:200]))
print(f"seeds {first}..{first + count - 1}: {count - len(bad)} passed, {len(bad)} failed")
for k, v in sorted(kinds.items()): print(f"  {v:4d} handles ran {k}")
for b in bad: print("  FAILED", b)
