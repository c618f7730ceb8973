"""B200-native ORB front-end: drop-in for ORB_SLAM2_ROS2's ORBExtractor + stereo / RGB-D depth association."""
__version__ = "0.1.0"
