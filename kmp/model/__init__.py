from kmp_demos_b200.model import KMP
