"""Drop-in ``kmp`` package: the reference's import paths (kmp/__init__.py:1, main.py:12) served by
kmp_demos_b200.  ``demo1/2/3`` resolve (lazily) to the headless variants in kmp_demos_b200.demos: the same
learn-and-adapt pipelines as the reference's matplotlib front-ends, with the arrays kept instead of drawn
(SURVEY.md section 8, row f4).  The unmodified reference demos.py also runs on top of this package."""


def __getattr__(name):
    if name in ('demo1', 'demo2', 'demo3'):
        from kmp_demos_b200 import demos
        return getattr(demos, name)
    raise AttributeError(f'module {__name__!r} has no attribute {name!r}')
