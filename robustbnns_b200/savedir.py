"""Path constants with the reference's names and values (savedir.py:4-6)."""
import time

DATA = "data/"
PLOTS = "plots/"
TESTS = "tests/" + str(time.strftime('%Y-%m-%d')) + "/"
