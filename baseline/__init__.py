"""Harness for running the UNMODIFIED reference (staged under baseline/_ref/, git-ignored) beside the product."""
