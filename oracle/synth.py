"""The seeded synthetic-input generator lives at the repo root (``synth.py``): it produces DATA only (weights with the
reference's parameter names, poses, latents) and is shared by the product-side bench arm, the tests and this oracle.
Re-exported here so that oracle-side code keeps one import."""
from synth import *  # noqa: F401,F403
