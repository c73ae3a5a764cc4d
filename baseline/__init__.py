"""Benchmark baselines (not product code; never imported by handobjectconsist_b200)."""
