"""the catkin shells (a-lego-loam_b200/ros/) keep the reference's plugin names, topics and cloud_info wire
format.  ROS is not in the build image, so the nodelet source is compiled against the stand-in headers of tests/ros_stubs/ (type
check of every call into alego_host.h and of every message field it touches); the real build is the CMakeLists.txt next to it."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROS = os.path.join(ROOT, "a-lego-loam_b200", "ros")


def test_nodelet_shells_compile_against_stub_headers(tmp_path):
    src = os.path.join(ROS, "src", "alego_nodelets.cpp")
    obj = str(tmp_path / "alego_nodelets.o")
    cmd = ["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fPIC", "-c", src, "-o", obj, "-I", os.path.join(ROOT, "tests", "ros_stubs"),
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "a-lego-loam_b200", "host")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    # the three plugin classes are instantiated (PLUGINLIB_EXPORT_CLASS) and reach the C ABI only through the host cores
    syms = subprocess.run(["nm", "-C", obj], capture_output=True, text=True).stdout
    for cls in ("loam::ImageProjection", "loam::LaserOdometry", "loam::LaserMapping"):
        assert cls + "::onInit()" in syms, cls
    for core in ("alego::ImageProjection::process", "alego::LaserOdometry::process", "alego::LaserMapping::process",
                 "alego::LaserOdometry::imuHandler", "alego::LaserMapping::extractSurroundingKeyFrames"):
        assert core in syms, core


def test_plugin_names_topics_and_message_match_the_reference_surface():
    xml = open(os.path.join(ROS, "nodelet_plugins.xml")).read()
    # reference nodelet_plugins.xml:1-11
    for name, typ in (("loam/ImageProjection", "loam::ImageProjection"), ("loam/LaserOdometry", "loam::LaserOdometry"),
                      ("loam/LaserMapping", "loam::LaserMapping")):
        assert re.search(r'class name="%s" type="%s" base_class_type="nodelet::Nodelet"' % (name, typ), xml)
    src = open(os.path.join(ROS, "src", "alego_nodelets.cpp")).read()
    # imageProjection.cpp:42-45, laserOdometry.cpp:52-72, laserMapping.cpp:82-93
    for topic in ("/lslidar_point_cloud", "/segmented_cloud", "/seg_info", "/outlier", "/imu/data", "/odom/lidar", "/surf_last",
                  "/corner_last", "/odom_aft_mapped"):
        assert '"%s"' % topic in src, topic
    # msg/cloud_info.msg:1-12 — and its C mirror AlegoCloudInfo
    msg = [l.split("#")[0].split() for l in open(os.path.join(ROS, "msg", "cloud_info.msg")) if l.split("#")[0].strip()]
    assert msg == [["Header", "header"], ["int32[]", "startRingIndex"], ["int32[]", "endRingIndex"], ["float32", "startOrientation"],
                   ["float32", "endOrientation"], ["float32", "orientationDiff"], ["bool[]", "segmentedCloudGroundFlag"],
                   ["int32[]", "segmentedCloudColInd"], ["float32[]", "segmentedCloudRange"]]
    hdr = open(os.path.join(ROOT, "include", "alego_b200.h")).read()
    body = hdr[hdr.index("typedef struct AlegoCloudInfo"):hdr.index("} AlegoCloudInfo;")]
    for _, field in msg[1:]:
        assert field in body, field
