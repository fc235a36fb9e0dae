"""State order of the DNA alphabet (reference: splitp/constants.py:7-8): fixes the 2-bit code."""
DNA_state_space = ("A", "C", "G", "T")
DNA_state_space_dict = {state: index for index, state in enumerate(DNA_state_space)}
