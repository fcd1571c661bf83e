"""dask_array_b200 -- B200-native execution backend for dask-array's data-parallel hot path."""
