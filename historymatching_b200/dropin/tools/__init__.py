"""B200-native drop-in for the notebooks' ``tools`` helper package
(reference ``notebooks/tools/``): ``utils`` (ensemble map, analysis
primitives), ``geostat`` (prior fields), ``localization`` (distances, taper),
``enopt`` (ensemble optimisation driver).  ``plotting`` is out of scope.
"""
