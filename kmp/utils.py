from kmp_demos_b200.utils import read_struct, find_closest_index, create_dataset
