"""Seeded synthetic inputs for tests, smoke() and bench.py -- not part of the product package."""
