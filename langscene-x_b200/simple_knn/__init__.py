"""Drop-in for the reference's `simple_knn` package (simple-knn/ext.cpp:15-17): `from simple_knn._C import distCUDA2`."""
