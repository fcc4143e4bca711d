// Empty stand-in: the reference's elements.h includes <ros/ros.h> but uses nothing from it.
#pragma once
