"""Test infrastructure: CPU oracle of the TA3N hot path (see ta3n_oracle.py).

Nothing in the product package ``ta3n_b200`` may import from here.
"""
