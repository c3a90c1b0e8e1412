#!/usr/bin/env python
"""Headline figures of one bench.py JSON line read from stdin (M solves/s):
value, single_stream, per_launch_flushed, e2e.   usage: python bench.py --no-extras | python tools/bench_line.py"""
import json
import sys

b = json.loads(sys.stdin.readlines()[-1])
print(round(b["value"] / 1e6, 1), round(b["single_stream"]["value"] / 1e6, 1),
      round(b["per_launch_flushed"]["value"] / 1e6, 1), round(b["e2e"]["value"] / 1e6, 1))
