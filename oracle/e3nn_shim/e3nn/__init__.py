"""TEST INFRASTRUCTURE ONLY.  Metadata-only stand-in for the `e3nn` package so the
reference file /root/reference/models/segnn/l1_tensor_prod.py imports unmodified
in this container (e3nn is not installed and there is no network).  Never
imported by the product package."""
