from kmp_demos_b200.cluster import Uniform, KMeans
