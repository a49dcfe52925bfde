#pragma once
#define GDL_B200_VERSION 100
