"""Drop-in ``kmp`` package: the reference's import paths (kmp/__init__.py:1, main.py:12) served by
kmp_demos_b200.  ``demo1/2/3`` are the reference's matplotlib front-ends and are out of scope here
(SURVEY.md section 2, rows 7-8); the unmodified reference demos.py runs on top of this package."""
