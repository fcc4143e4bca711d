// Empty stand-in: the reference's elements.h includes <geometry_msgs/Pose.h> but uses nothing from it.
#pragma once
