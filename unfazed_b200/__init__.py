__version__ = "1.0.3+b200.r1"
