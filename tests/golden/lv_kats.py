"""Known-answer tests of the reference, restated as data.

Source: tests/LandauVishkinTest.cpp:11-32 (edit distance, 11 cases) and :34-130 (CIGAR, 36 cases) of
andrewmagis/snap-rnaseq.  Tuple layouts:
    SCORE_KATS: (text, pattern, k, expected distance)
    CIGAR_KATS: (text, pattern, k, useM, expected CIGAR)
"""
SCORE_KATS = [
    ("abcde", "abcde", 2, 0),
    ("abcde", "abcd", 2, 0), ("abcde", "abc", 2, 0), ("abcde", "ab", 2, 0),
    ("abcde", "abcdX", 2, 1), ("abcde", "abde", 2, 1), ("abcde", "bcde", 2, 1), ("abcde", "abcXde", 2, 1),
    ("abcde", "abXXe", 2, 2), ("abcde", "abcXXde", 2, 2),
    ("abcde", "XXXXX", 2, -1),
]

CIGAR_KATS = [
    ("abcde", "abcde", 2, False, "5="), ("abcde", "abcde", 2, True, "5M"),
    ("abcdef", "abcde", 2, False, "5="), ("abcdef", "abcde", 2, True, "5M"),
    ("abcde", "abcdX", 2, False, "4=1X"), ("abcde", "abcdX", 2, True, "5M"),
    ("abcde", "Xbcde", 2, False, "1X4="), ("abcde", "Xbcde", 2, True, "5M"),
    ("abcde", "abde", 2, False, "2=1D2="), ("abcde", "abde", 2, True, "2M1D2M"),
    ("abcde", "bcde", 2, False, "1D4="), ("abcde", "bcde", 2, True, "1D4M"),
    ("abcde", "abcXde", 2, False, "3=1I2="), ("abcde", "abcXde", 2, True, "3M1I2M"),
    ("abcde", "abXXe", 2, False, "2=2X1="), ("abcde", "abXXe", 2, True, "5M"),
    ("abcde", "abcXXde", 3, False, "3=2I2="), ("abcde", "abcXXde", 3, True, "3M2I2M"),
    ("ttttc", "tttc", 3, False, "3=1X"), ("ttttc", "tttc", 3, True, "4M"),
    ("tttcc", "ttttc", 3, False, "3=1X1="), ("tttcc", "ttttc", 3, True, "5M"),
    ("tttcc", "tttaa", 3, False, "3=2X"), ("tttcc", "tttaa", 3, True, "5M"),
    ("atctcag", "acttcag", 3, False, "1=2X4="), ("atctcag", "acttcag", 3, True, "7M"),
    ("abc", "abcde", 3, False, "3=2X"), ("abc", "abcde", 3, True, "5M"),
    ("abc", "abXde", 3, False, "2=3X"), ("abc", "abXde", 3, True, "5M"),
]

# ProbabilityDistance KATs, tests/ProbabilityDistanceTest.cpp:15-70 of the reference (fixture: ProbabilityDistance(0.1, 0.01, 0.2);
# ASSERT_NEAR there is +-1 %, tests/TestLib.h:136-141): (reference, read, quality, maxStartShift, maxShift, expected probability)
Q10 = chr(43)
PROBABILITY_DISTANCE_KATS = [
    ("A", "A", "I", 0, 0, 0.9),
    ("A", "C", "I", 0, 0, 0.1),
    ("A", "C", Q10, 0, 0, 0.19),
    ("A", "A", "I", 1, 2, 0.9),
    ("A", "C", "I", 1, 2, 0.1),
    ("A", "C", Q10, 1, 2, 0.19),
    ("AAAAA", "AAAAA", "IIIII", 1, 2, 0.9 ** 5),
    ("AAAAA", "AACAA", "IIIII", 1, 2, 0.9 ** 4 * 0.1),
    ("ACGTA", "ACGGTA", "IIIIII", 1, 2, 0.9 ** 5 * 0.01),
    ("ACGTA", "ACTA", "IIII", 1, 2, 0.9 ** 2 * 0.1 ** 2),
    ("ACGTACGT", "ACGTTACGT", "IIIIIIIII", 1, 2, 0.9 ** 8 * 0.01),
    ("ACGTACGT", "ACGACGT", "IIIIIII", 1, 2, 0.9 ** 7 * 0.01),
    ("ACGTACGT", "ACTACGT", "IIIIIII", 0, 2, 0.9 ** 7 * 0.01),
    ("ACGTACGT", "ACTACGT", "IIIIIII", 1, 2, 0.9 ** 5 * 0.1 ** 2),
    ("ACGTACGT", "ACGTTTACGT", "IIIIIIIIII", 1, 2, 0.9 ** 8 * 0.01 * 0.2),
    ("ACGTTTACGT", "ACGTACGT", "IIIIIIII", 1, 2, 0.9 ** 8 * 0.01 * 0.2),
]
