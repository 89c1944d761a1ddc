"""Three-symbol stand-in for `timm` so the UNMODIFIED reference module
(/root/reference/models/hit_sir_pro.py:6) can be imported in this container
(timm is not installed, no network).  Test/golden-generation infrastructure only."""
