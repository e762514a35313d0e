// the reference includes this header and never calls it (PSC/main.cpp:15)
#pragma once
