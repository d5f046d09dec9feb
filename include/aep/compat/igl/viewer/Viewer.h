/* igl/viewer/Viewer.h for machines without libigl / OpenGL: the headless viewer (include/aep/headless). */
#pragma once
#include "../../../headless/igl/viewer/Viewer.h"
