__version__ = "0.2.1+b200.r1"      # mirrors casingSimulations/info.py:1 (0.2.1)
__author__ = "casing-b200 builders"
__license__ = "MIT"
