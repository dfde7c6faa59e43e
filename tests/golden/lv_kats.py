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
