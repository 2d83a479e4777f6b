"""Golden-vector case table shared by make_golden.py (generator) and the tests."""
# name, B, Lt, Li, R, reversed, training, realistic, param_scale
CASES = [
    ("text_r3_train", 4, 12, 7, 3, False, True, False, 1.0),
    ("image_r3_train", 4, 12, 7, 3, True, True, False, 1.0),
    ("text_r4_eval", 3, 16, 5, 4, False, False, False, 1.0),
    ("image_r3_eval_real", 2, 9, 10, 3, True, False, True, 1.0),
    ("text_r3_train_real", 3, 10, 6, 3, False, True, True, 1.0),
]
PARAM_SEED_BASE = 2023   # params seed = PARAM_SEED_BASE + R
INPUT_SEED_BASE = 7      # inputs seed = INPUT_SEED_BASE + B
LOSS_SEED = 99           # loss = sum(out * w_out) + sum(sim * w_sim), w ~ N(0,1) from this seed
