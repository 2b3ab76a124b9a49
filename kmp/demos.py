"""kmp.demos -> kmp_demos_b200.demos (headless demo1 / demo2 / demo3)."""
from kmp_demos_b200.demos import demo1, demo2, demo3, main  # noqa: F401
