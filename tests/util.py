"""Shared helpers for the test-suite."""
import gzip
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(GOLDEN, "data")


def data_path(name):
    return os.path.join(DATA, name)


def golden_cli_output(name):
    with gzip.open(os.path.join(GOLDEN, "cli", name + ".out.gz"), "rb") as fh:
        return fh.read()
