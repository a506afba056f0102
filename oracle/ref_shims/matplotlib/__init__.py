"""Placeholder for matplotlib (uf3.forcefield.properties.phonon imports it at module load)."""
