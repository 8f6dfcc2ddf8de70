// TEST INFRASTRUCTURE ONLY (oracle/): see App.hpp
#pragma once
