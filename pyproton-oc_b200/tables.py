"""The three reference tables the crime path reads, as the reference itself loads them.

Sources (reference root): protonoc/simulator/inputs/palermo/data/crime_rate_by_gender_and_age_range.csv
(model.py:177-178), inputs/general/data/num_co_offenders_dist.csv (model.py:188-190),
inputs/palermo/data/conviction_length.csv (model.py:170-174).

The reference reads them with pandas.read_csv's default float parser, which keeps 15 significant digits
("0.00220888483881099" -> 0.0022088848388109), so these are the values its arithmetic actually uses; they
were dumped with repr() from a loaded reference model (oracle/ref_harness.py).
"""

# (male, age_from, age_to, c) in file order
C_RANGE_BY_AGE_AND_SEX = [
    (0, 0, 13, 0.000419095530846447), (1, 0, 13, 0.0022088848388109),
    (0, 14, 17, 0.0223485080056817), (1, 14, 17, 0.15015920337183),
    (0, 18, 24, 0.0510827740184912), (1, 18, 24, 0.301852378661966),
    (0, 25, 34, 0.0634378338266472), (1, 25, 34, 0.303561858487422),
    (0, 35, 44, 0.0642953324827881), (1, 35, 44, 0.275109864781256),
    (0, 45, 54, 0.0488535842667994), (1, 45, 54, 0.19961940993774),
    (0, 55, 64, 0.0308128546507127), (1, 55, 64, 0.126843825050472),
    (0, 65, 200, 0.0111212661509359), (1, 65, 200, 0.0536705333692047)]

# [n, p]
NUM_CO_OFFENDERS_DIST = [[1, 0.8744024339984563], [2, 0.0960396945979063], [3, 0.0207191652676337],
                         [4, 0.0055633579042241], [5, 0.0018980490549422], [6, 0.0013772991768372]]

# (months, F, M)
CONVICTION_LENGTH = [
    (1, 0.135167397354478, 0.078947866071776), (3, 0.19269400131653, 0.131143083933566),
    (6, 0.262419887899263, 0.236163183846043), (12, 0.205487284788606, 0.238697628478702),
    (24, 0.140485495612569, 0.175982293376721), (36, 0.032234256152626, 0.059220489881712),
    (60, 0.021495330654007, 0.049035337226683), (120, 0.008189481085251, 0.022837548624071),
    (240, 0.00182686513667, 0.007972568560726)]
