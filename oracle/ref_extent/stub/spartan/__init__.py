# Import stub so the reference's extent.pyx (which does `from spartan import util`) can load
# stand-alone.  Only the three helpers extent.pyx touches are provided.
