"""Config keys of the hot path with the reference's names and defaults (core/config.py:8-28,32-37,57-66,76-80,100-106).
Only these keys are read; any reference `myconfig.json` dict can be passed as-is."""

_DEFAULTS = {
    'img_w': 160,
    'img_h': 120,
    'cam_resolution': [320, 240],
    'preprocessing_enabled': False,
    'preprocessing_preview_enabled': True,   # accepted, ignored: there is no imshow on a GPU box
    'preprocessing_contrast_enhancement_ratio': 1.0,
    'preprocessing_contrast_enhancement_offset': 125,
    'preprocessing_dynamic_brightness_enabled': False,
    'preprocessing_brightness_baseline': 550,
    'preprocessing_color_filter_enabled': False,
    'preprocessing_color_filter_hsvs': [((0, 0, 130), (180, 64, 255)), ((25, 180, 155), (43, 255, 255))],
    'preprocessing_color_filter_destination_channels': [0, 1],
    'preprocessing_edge_detection_enabled': False,
    'preprocessing_edge_detection_threshold_a': 60,
    'preprocessing_edge_detection_threshold_b': 100,
    'preprocessing_edge_detection_destination_channel': 2,
    'smooth_steering_enabled': False,
    'smooth_steering_threshold': 0.9,
    'spd_ctl_threshold': 1.1,
    'spd_ctl_reverse': True,
    'spd_ctl_reverse_multiplier': 1.0,
    'spd_ctl_break': False,
    'spd_ctl_break_multiplier': 1.0,
    # per-car control post-processing (controlmultiplexer.py:11-20, driver_assistance.py:10-11, teensy.py PWM calibration)
    'calibrate_max_forward_pwm': 400,
    'calibrate_zero_throttle_pwm': 370,
    'calibrate_max_reverse_pwm': 330,
    'calibrate_max_left_pwm': 430,
    'calibrate_max_right_pwm': 300,
    'calibrate_neutral_steering_pwm': 350,
    'ai_launch_boost_throttle_enabled': False,
    'ai_launch_boost_throttle_value': 1.0,
    'ai_launch_boost_throttle_duration': 5,
    'ai_launch_lock_steering_enabled': False,
    'ai_launch_lock_steering_value': 0.0,
    'ai_launch_lock_steering_duration': 3,
    'drive_assist_enabled': False,
    'drive_assist_limit_mode': 'steering',
    'drive_assist_limit_k': 5,
    'use_location_tracker': False,
    'track_data_file': 'track_data/generated_track.json',
}


def default_config(**overrides):
    cfg = {k: (list(v) if isinstance(v, list) else v) for k, v in _DEFAULTS.items()}
    cfg.update(overrides)
    return cfg


def full_house_config(**overrides):
    """cnn_2d_full_house observation: both colour ranges + edge mask (BASELINE.json configs[2])."""
    return default_config(preprocessing_enabled=True, preprocessing_color_filter_enabled=True,
                          preprocessing_edge_detection_enabled=True, **overrides)
