"""Drop-in for `simple_knn._C` (submodules/simple-knn/ext.cpp:15-17):
`from simple_knn._C import distCUDA2` (scene/gaussian_model.py:20)."""
from gftorf_b200.knn import distCUDA2  # noqa: F401
