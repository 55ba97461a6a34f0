"""CPU stand-in (tests only): the product's distCUDA2 takes a brute-force path for tiny CPU clouds."""
from instantsplat_b200.knn import distCUDA2  # noqa: F401
