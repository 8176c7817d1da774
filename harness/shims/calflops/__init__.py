"""Stand-in for `calflops` (absent): the reference's utils.py:4 imports it at module level for print_flops_params only."""


def calculate_flops(*args, **kwargs):
    raise NotImplementedError("calflops is not installed")
