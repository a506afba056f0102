"""Placeholder for PyTables: only uf3.data.io's HDF5 helpers need it (not on the hot path)."""
