"""Fixture loading helpers shared by the CPU and GPU suites."""
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(GOLDEN, name)) as fh:
        return json.load(fh)


def ints(v):
    return [int(x) for x in v]


def g1(p):
    return None if p is None else (int(p[0]), int(p[1]))


def g2(p):
    return None if p is None else ((int(p[0][0]), int(p[0][1])), (int(p[1][0]), int(p[1][1])))
