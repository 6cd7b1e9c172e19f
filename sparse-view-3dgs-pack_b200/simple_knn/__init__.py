"""Drop-in `simple_knn` package: `from simple_knn._C import distCUDA2` (LG/scene/gaussian_model.py:21,159)."""
